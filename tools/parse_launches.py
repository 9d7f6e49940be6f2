"""Summarise an ncu --metrics gpu__time_duration.sum launch list of tools/profile_step.py (last call, first step)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
iK, iV, iG = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
data = [(r[iK][:34], r[iG], float(r[iV].replace(',', ''))) for r in rows[hdr + 2:] if len(r) > iV]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 16
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
last = data[-per * K:]
tot = sum(d[2] for d in last) / K
names = ["stage_z", "CT1 fwd", "CT2 fwd c0", "CT2 fwd c1", "CT2 fwd c2", "CT2 fwd c3", "CT3 fwd c0", "CT3 fwd c1",
         "CT3 fwd c2", "CT3 fwd c3", "CT4 fwd(last)", "CT4 dgrad(col)", "CT3 dgrad", "CT2 dgrad", "CT1 dgrad(z)",
         "ebm tail"]
for i, d in enumerate(last[:per]):
    nm = names[i] if per == 16 else str(i)
    print(f"{nm:16s} {d[0]:34s} grid {d[1]:>12s} {d[2] / 1e3:9.1f} us  {100 * d[2] / tot:5.1f}%")
print("per-step total us %.1f" % (tot / 1e3))
