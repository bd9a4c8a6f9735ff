"""Build the sm_100a C-ABI library in-tree: multiviewhmr_b200/lib/libmvhmr_b200.so.

Explicit nvcc (cross-compiles without a GPU); the built .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libmvhmr_b200.so")
SOURCES = ["abi.cu", "geometry.cu", "unproject.cu", "unproject_out0.cu", "unproject_out1.cu", "unproject_out2.cu", "unproject_out3.cu",
           "unproject_staged.cu", "unproject_tex.cu", "softargmax.cu", "backward.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
LINK_FLAGS = ["-shared", "-cudart", "static"]
OBJ_DIR = os.path.join(HERE, "build")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", "mvhmr_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build(force=False, verbose=False):
    """Every translation unit is compiled on its own (in parallel), then linked into the .so."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    jobs = [(nvcc, os.path.join(CSRC, s), os.path.join(OBJ_DIR, s.replace(".cu", ".o")), verbose) for s in SOURCES]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
        results = list(pool.map(_compile, jobs))
    for src, rc, log in results:
        if rc != 0:
            sys.stderr.write(log)
            raise RuntimeError("nvcc failed compiling %s" % src)
        if verbose:
            sys.stderr.write(log)
    cmd = [nvcc] + NVCC_FLAGS + LINK_FLAGS + ["-o", LIB] + [j[2] for j in jobs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking %s" % LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
