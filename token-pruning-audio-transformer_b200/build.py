"""Build libtpat.so (sm_100a only) in-tree with nvcc.

    python token-pruning-audio-transformer_b200/build.py [--force]

Each csrc/*.cu is compiled to an object (in parallel, rebuilt only when the source, a header or
the flags changed) and linked into ``lib/libtpat.so``.  nvcc cross-compiles without a GPU, so this
runs in the build container; the .so travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtpat.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())     # not the full path: the repo lives elsewhere on the GPU box
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    sources = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(INCLUDE, "tpat.h"))
    hdr_digest = _digest(headers, " ".join(FLAGS))
    # library-level stamp: the .so shipped to the GPU box (the object directory does not travel) is reused as is when
    # no source, header or flag changed
    lib_stamp = LIB_PATH + ".sha"
    lib_want = _digest(sources, hdr_digest)
    if not force and os.path.exists(LIB_PATH) and os.path.exists(lib_stamp) and open(lib_stamp).read() == lib_want:
        if verbose:
            print(f"[tpat build] {LIB_PATH} (up to date; {len(sources)} sources)")
        return LIB_PATH

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        want = _digest([src], hdr_digest)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
            return obj, False
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(want)
        return obj, True

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    objs = [o for o, _ in results]
    rebuilt = any(c for _, c in results)
    if rebuilt or force or not os.path.exists(LIB_PATH):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(lib_stamp, "w") as f:
        f.write(lib_want)
    if verbose:
        print(f"[tpat build] {LIB_PATH} ({'rebuilt' if rebuilt else 'up to date'}; {len(objs)} objects)")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
