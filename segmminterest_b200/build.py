"""Builds libmmi_b200.so (in-tree) with nvcc for sm_100a.  No JIT cache: the .so sits next
to the sources so it travels to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmmi_b200.so")
SOURCES = ["api.cu", "gather.cu", "gemm_simt.cu", "gemm_tc.cu", "tc_host.cu", "elementwise.cu", "attention_simt.cu", "attention_tc.cu", "idfusion.cu", "metrics.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    if os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "mmi_b200.h")) > t:
        return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "-Xcompiler", "-fPIC", "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        cmd[1:1] = os.environ.get("MMI_NVCC_EXTRA", "").split()
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {s} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {s} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc build of libmmi_b200.so failed")
    cmd = [NVCC, "-shared", "-o", LIB] + objs  # static cudart (nvcc default); driver entry points via cudaGetDriverEntryPoint
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
