"""In-tree build of the engine's shared library (nvcc cross-compiles sm_100a without a GPU).

    python stable-neural-sdes_b200/build.py [--force] [--verbose]

Each csrc/*.cu is compiled to an object (cached by content hash of the unit + all headers, in parallel) and
linked into ``stable-neural-sdes_b200/libsnsde.so``; objects and the library are git-ignored but travel to the
GPU box with the repo snapshot.
"""
import concurrent.futures
import hashlib
import pathlib
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "libsnsde.so"
import os

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
if os.environ.get("SNSDE_TRACE_BUILD"):       # debug build: compiles the clock64 trace events into the tcgen05 kernels;
    NVCC_FLAGS.append("-DSNSDE_TC_TRACE_BUILD")   # a separate library (loaded when SNSDE_TRACE_BUILD is set), own object cache
    OBJ = HERE / "build_trace"
    LIB = HERE / "libsnsde_trace.so"


def _headers_hash():
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "snsde.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src, hdr_hash, force, verbose):
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".stamp")
    fp = hashlib.sha256(src.read_bytes() + hdr_hash.encode()).hexdigest()
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == fp:
        return obj, False, ""
    cmd = ["nvcc", *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", str(obj), str(src)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{res.stdout}{res.stderr}")
    stamp.write_text(fp)
    return obj, True, res.stdout + res.stderr


def build(force=False, verbose=False):
    hdr_hash = _headers_hash()
    sources = sorted(CSRC.glob("*.cu"))
    # whole-library fingerprint next to the .so: lets a snapshot without the object cache (GPU box) skip the build
    lib_fp = hashlib.sha256("".join(hashlib.sha256(s.read_bytes()).hexdigest() for s in sources).encode()
                            + hdr_hash.encode()).hexdigest()
    lib_stamp = HERE / ("." + LIB.stem + ".stamp")
    if not force and LIB.exists() and lib_stamp.exists() and lib_stamp.read_text() == lib_fp:
        return LIB
    OBJ.mkdir(exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda s: _compile(s, hdr_hash, force, verbose), sources))
    if verbose:
        for _, _, log in results:
            sys.stderr.write(log)
    if any(changed for _, changed, _ in results) or not LIB.exists():
        # cuBLAS: the weight-gradient GEMMs of the backward pass (plain library GEMMs); resolved at load time to the
        # copy torch has already loaded (same soname), or to the toolkit's through the rpath
        cmd = ["nvcc", "-shared", "-o", str(LIB), *[str(o) for o, _, _ in results], "-lcublas",
               "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    lib_stamp.write_text(lib_fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
