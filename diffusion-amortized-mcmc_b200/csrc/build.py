"""Build libdamc_b200.so in-tree with nvcc for sm_100a (the .so travels to the GPU box with the repo snapshot)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "damc_b200", "libdamc_b200.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "8",
         "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v" if os.environ.get("DAMC_PTXAS_V") else "-O3"]


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(HERE, "*.cu")))
    deps = srcs + glob.glob(os.path.join(HERE, "*.h")) + glob.glob(os.path.join(HERE, "*.cuh")) + \
        [os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "damc.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["nvcc"] + FLAGS + ["-o", OUT] + srcs + ["-lcudart"]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdamc_b200.so")
    if verbose or os.environ.get("DAMC_PTXAS_V"):
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose=True))
