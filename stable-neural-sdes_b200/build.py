"""In-tree build of the engine's shared library (nvcc cross-compiles sm_100a without a GPU).

    python stable-neural-sdes_b200/build.py [--force] [--verbose]

Produces ``stable-neural-sdes_b200/libsnsde.so`` next to this file; it is git-ignored but
travels to the GPU box with the repo snapshot.
"""
import hashlib
import pathlib
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libsnsde.so"
STAMP = HERE / ".libsnsde.stamp"
SOURCES = ["snsde_api.cu", "snsde_fma.cu", "snsde_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _fingerprint():
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "snsde.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    fp = _fingerprint()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text() == fp:
        return LIB
    cmd = ["nvcc", *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", str(LIB),
           *[str(CSRC / s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsnsde.so")
    STAMP.write_text(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
