"""Print the last posterior step of an ncu launch list (csv from tools/gpu_r2_*.sh): time, DRAM bytes, tensor-pipe activity."""
import sys
sys.path.insert(0, 'tools')
import summarize_launches as s
L = s.load(sys.argv[1])
idx = [i for i, e in enumerate(L) if 'stage_z' in e['k']]
L = L[idx[-1]:]
L = [e for e in L if not e['k'].startswith('at::')]
tot = sum(e["gpu__time_duration.sum"] for e in L)
print(sys.argv[1])
for e in L:
    t = e["gpu__time_duration.sum"]; rd, wr = e.get("dram__bytes_read.sum", 0), e.get("dram__bytes_write.sum", 0)
    tp = e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)
    print(f"{e['k'][:40]:40s} grid {e['grid']:>12s} {t:8.1f} us {100*t/tot:5.1f}%  rd {rd:7.1f} wr {wr:7.1f} MB ({(rd+wr)/t:5.2f} TB/s) tp {tp:5.1f}%")
print(f"total {tot:.1f} us")
