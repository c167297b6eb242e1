"""The ctypes mirrors in segmminterest_b200/_lib.py must have the memory layout a C compiler gives the structs of
include/mmi_b200.h: a C translation unit that includes the header prints sizeof / offsetof, ctypes has to agree.
(CPU only: gcc compiles the header as plain C, which also proves the ABI has no C++ in it.)"""
import ctypes as C
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FIELDS = {
    "mmi_dropout": ["key", "thr8", "scale"],
    "mmi_gemm_args": ["layout", "M", "A", "lda", "C", "bias", "act", "preact", "mul_gelu_grad", "add", "add_mod", "add_dtype", "accumulate",
                      "split_k", "save_act_grad", "mul_is_grad", "drop", "mul_scale", "split_ws", "split_ws_bytes"],
    "mmi_attn_block": ["q", "ldq", "mask_k", "Lk", "dq", "lddv", "dbq", "dbv"],
    "mmi_attn_args": ["dtype", "Lq", "mask_q", "nblk", "blk", "out", "lse", "dout", "delta", "drop", "dq_acc", "dq_count"],
    "mmi_loss_args": ["logits", "gt", "B", "exposure_prob", "inv_bsz", "use_focal", "logits_out", "dbias_bias", "use_huber", "kl_after_focal",
                      "w_huber", "w_interestKL"],
}


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_ctypes_structs_match_the_c_header(tmp_path):
    from segmminterest_b200 import _lib
    mirrors = {"mmi_dropout": _lib.Dropout, "mmi_gemm_args": _lib.GemmArgs, "mmi_attn_block": _lib.AttnBlock,
               "mmi_attn_args": _lib.AttnArgs, "mmi_loss_args": _lib.LossArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mmi_b200.h"', "int main(void) {"]
    for st, fields in FIELDS.items():
        lines.append(f'  printf("{st} sizeof %zu\\n", sizeof({st}));')
        for f in fields:
            lines.append(f'  printf("{st} {f} %zu\\n", offsetof({st}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True)
    seen = 0
    for line in out.strip().splitlines():
        st, what, val = line.split()
        m = mirrors[st]
        want = C.sizeof(m) if what == "sizeof" else getattr(m, what).offset
        assert int(val) == want, f"{st}.{what}: C says {val}, ctypes says {want}"
        seen += 1
    assert seen == sum(len(v) + 1 for v in FIELDS.values())
