"""Build libmau_b200.so in-tree with nvcc for sm_100a (no torch headers, no libcuda link) and the host-only
libmau_tiles.so (tile reader, g++ + zlib)."""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(SRC, "obj")
LIB = os.path.join(HERE, "libmau_b200.so")
TILES_SRC = os.path.join(SRC, "tile_reader.cpp")
TILES_LIB = os.path.join(HERE, "libmau_tiles.so")
CXX = os.environ.get("CXX", "g++")
SOURCES = ["common.cu", "tma.cu", "conv_tc.cu", "wgrad_tc.cu", "conv_ffma.cu", "elementwise.cu", "norm.cu",
           "encoders.cu", "loss.cu", "ssim.cu", "metrics.cu", "optim.cu", "embgrad.cu", "plan.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
         "-Xcompiler", "-fPIC", "-DNDEBUG"]


def _newer(src, dst, deps):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "mau_b200.h"))
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(SRC, s), os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _newer(src, obj, headers):
            jobs.append((src, obj))

    def cc(job):
        cmd = [NVCC, *FLAGS, "-c", job[0], "-o", job[1]]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {job[0]}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr)

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        list(ex.map(cc, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_tiles(force=force)
    return LIB


def build_tiles(force=False):
    """Host-only tile reader (include/mau_tiles.h): g++ + zlib, no CUDA."""
    hdr = os.path.join(HERE, "..", "include", "mau_tiles.h")
    if force or _newer(TILES_SRC, TILES_LIB, [hdr, os.path.join(SRC, "inflate_fast.h")]):
        cmd = [CXX, "-O3", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-DNDEBUG", TILES_SRC, "-o", TILES_LIB,
               "-lz", "-pthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {TILES_SRC}:\n{r.stdout}\n{r.stderr}")
        if r.stderr.strip():
            print(r.stderr, file=sys.stderr)
    return TILES_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
