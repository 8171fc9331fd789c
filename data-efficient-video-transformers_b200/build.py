"""Build libtvt_b200.so (hand-written sm_100a kernels + the C-ABI) in-tree with nvcc.

    python "data-efficient-video-transformers_b200/build.py" [--force] [-v]

nvcc cross-compiles without a GPU; the shared object is git-ignored but travels to the GPU box.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libtvt_b200.so")
SOURCES = ["core.cu", "gemm_sm100.cu", "layernorm.cu", "attention_simt.cu", "attention_sm100.cu", "pool.cu", "loss.cu", "misc.cu", "collab.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _headers():
    inc = os.path.join(os.path.dirname(HERE), "include", "tvt.h")
    return [inc] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    stamp = os.path.join(OBJ, "stamp.txt")
    want = _digest(srcs + _headers())
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIB
    hdr_digest = _digest(_headers())

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        tag = obj + ".tag"
        d = _digest([src]) + hdr_digest
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read() == d:
            return obj
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(tag, "w") as f:
            f.write(d)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
