"""Compiles the round-2 prototype (fast_lio_b200/csrc/experimental/cellmap.cu) into its own library.  It is NOT part of
libfastlio_b200.so, of __graft_entry__.build() or of the default test run."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "fast_lio_b200", "csrc", "experimental", "cellmap.cu")
OUT = os.path.join(ROOT, "fast_lio_b200", "libfastlio_b200_experimental.so")
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--fmad=false",
       "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-Xptxas", "-v", SRC, "-o", OUT + ".tmp"]
r = subprocess.run(cmd, capture_output=True, text=True)
sys.stderr.write("\n".join(l for l in (r.stdout + r.stderr).splitlines() if "k_cell_knn" in l or "registers" in l or "error" in l) + "\n")
if r.returncode != 0:
    sys.stderr.write(r.stdout + r.stderr)
    sys.exit(1)
os.replace(OUT + ".tmp", OUT)
print(OUT)
