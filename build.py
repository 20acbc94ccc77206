"""Build libjmt_b200.so (sm_100a) in-tree with nvcc, and the oracle checkers.

    python build.py            # build if sources are newer than the library
    python build.py --force
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "joint-multimodal-transformer-6th-abaw_b200")
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libjmt_b200.so")
SOURCES = ["core.cu", "ccc.cu", "rowops.cu", "elementwise.cu", "heads.cu", "gemm_simt.cu", "gemm_tc.cu", "attn_tc.cu", "attn_bwd_tc.cu", "valpost.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC",
              "-cudart", "static"]


def _newer(src_files, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in src_files)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(ROOT, "include", "jmt_b200.h")]
    if not force and not _newer(deps, LIB):
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(PKG, "build", os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["nvcc", *NVCC_FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    fail = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {os.path.basename(s)}\n{out}\n")
        fail |= p.returncode != 0
    if fail:
        raise RuntimeError("nvcc failed")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", LIB, *objs]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
