"""No GPU needed: the built library is Blackwell-native.  Its SASS must contain the tcgen05 / TMEM / TMA opcodes of the hot kernels
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store) and it must not link a BLAS."""
import os
import re
import shutil
import subprocess

import pytest

from mirror_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")
def test_library_contains_tcgen05_tmem_tma_opcodes():
    _lib.lib()
    sass = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    per_kernel = {}
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        for op in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG"):
            if cur and re.search(r"\b" + op + r"\b", line):
                per_kernel.setdefault(cur, set()).add(op)
    has = lambda frag, ops: any(frag in k and set(ops) <= v for k, v in per_kernel.items())
    assert has("gemm_tcgen05_kernel", ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG"))
    assert has("gemm_tcgen05_multi_kernel", ("UTCHMMA", "LDTM", "UTMALDG"))
    # the flash kernels keep probabilities / dS in tensor memory: they must STORE to it as well (the TS-form A operand)
    assert has("flash_fwd_kernel", ("UTCHMMA", "LDTM", "STTM", "UTMALDG"))
    assert has("flash_fwd_twopass_kernel", ("UTCHMMA", "LDTM", "STTM", "UTMALDG"))
    assert has("flash_bwd_kernel", ("UTCHMMA", "LDTM", "STTM", "UTMALDG"))
    assert has("contrastive_kernel", ("UTCHMMA", "LDTM", "UTMALDG"))
    assert "sm_100a" in sass


def test_library_links_no_blas_or_dnn():
    _lib.lib()
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "cublas" not in out.lower() and "cudnn" not in out.lower(), out
