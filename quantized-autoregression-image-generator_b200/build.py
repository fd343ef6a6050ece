"""Build libsomcb.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python quantized-autoregression-image-generator_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "somcb", "libsomcb.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["som_core.cu", "som_filter.cu", "som_filter_tc.cu", "som_accumulate.cu", "som_bmu_ffma.cu", "som_bmu_tc.cu",
           "som_bmu_tc_s.cu", "som_bmu_tc_l.cu", "som_bmu_tc_l16.cu", "som_bmu.cu", "som_peer.cu", "som_step_small.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
         "-Xptxas", "-v", "-DSOM_BUILDING_LIB"]
FLAGS += os.environ.get("SOMCB_EXTRA_NVCC_FLAGS", "").split()        # e.g. -DSOM_TC_EXPERIMENTS (tools/ only)


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(__file__))


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _deps_mtime():
        return OUT
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; libsomcb.so cannot be built")
    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "--cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
