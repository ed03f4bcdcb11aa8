"""Build libammc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m ammcnet_aaai2021_b200.build [--force]

The shared library has a plain C ABI (include/ammc_b200.h), links the CUDA runtime statically and takes
the driver's cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, so it loads on machines without a GPU.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libammc_b200.so")
STAMP = os.path.join(PKG, "csrc", ".build_stamp")
SOURCES = ["core.cu", "mem_simt.cu", "addr_tc.cu", "enc_tc.cu", "mem_front.cu", "score.cu", "auc.cu", "amft_conv.cu", "amft_train.cu", "unet_ops.cu", "halo_conv.cu", "preprocess.cu", "losses.cu"]
DEBUG_SOURCES = SOURCES + ["probes.cu"]            # + -DAMMC_DEBUG_PROBES -> libammc_b200_debug.so (tools only)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the sm_100a library cannot be built")


def _digest():
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    files = [os.path.join(CSRC, f) for f in files] + [os.path.join(PKG, "..", "include", "ammc_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True, debug: bool = False) -> str:
    """Product library (default) or, with debug=True, libammc_b200_debug.so = the same sources + the hardware probes
    (include/ammc_b200_debug.h), compiled with -DAMMC_DEBUG_PROBES into their own objects."""
    lib = os.path.join(PKG, "libammc_b200_debug.so") if debug else LIB
    stamp = STAMP + (".debug" if debug else "")
    digest = _digest()
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return lib
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in (DEBUG_SOURCES if debug else SOURCES):
        obj = os.path.join(CSRC, src.replace(".cu", ".dbg.o" if debug else ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-DAMMC_DEBUG_PROBES"] if debug else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out.decode(errors="replace")))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as fh:
        fh.write(digest)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, debug="--debug" in sys.argv))
