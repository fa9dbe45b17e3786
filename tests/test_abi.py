"""CPU: the C-ABI library loads, exports every symbol include/mirror_b200.h declares, and the ctypes
signature table used by the Python host side matches the header argument by argument.  No compute calls."""
import ctypes
import os
import re

from mirror_b200 import _lib

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mirror_b200.h")
CT = {"int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "uint64_t": ctypes.c_uint64}


def _decls():
    src = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    for m in re.finditer(r"\b(?:int|const char\*)\s+(mirror_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, re.S):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        yield m.group(1), args


def test_header_symbols_exported():
    lib = _lib.lib()
    names = [n for n, _ in _decls()]
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(names) == sorted(_lib.EXPORTS)


def test_signature_table_matches_header():
    for name, args in _decls():
        if name not in _lib.SIGNATURES:
            continue
        want = []
        for a in args:
            if "*" in a or a.startswith("mirror_stream_t"):
                want.append(ctypes.c_void_p)
            else:
                want.append(CT[a.split()[0]])
        assert want == _lib.SIGNATURES[name], f"{name}: header {args} vs table"


def test_gemm_struct_layout_matches_header():
    src = open(HDR).read()
    body = re.search(r"typedef struct \{(.*?)\} mirror_gemm_args;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        ptr = "*" in stmt
        ty = stmt.split()[1] if stmt.startswith("const") else stmt.split()[0]
        names = [x.strip().lstrip("*") for x in stmt.replace("*", " * ").split(None, 2 if stmt.startswith("const") else 1)[-1].replace("*", "").split(",")]
        for n in names:
            fields.append((n.strip(), ctypes.c_void_p if ptr else CT[ty]))
    assert [(n, t) for n, t in _lib.GemmArgs._fields_] == fields


def test_abi_version_and_error_string():
    lib = _lib.lib()
    assert lib.mirror_abi_version() == 1
    assert isinstance(lib.mirror_last_error(), bytes)


def test_kernels_refuse_cpu_tensors():
    import pytest
    import torch
    from mirror_b200 import kernels as K
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        K.gemm(a, a, out_f32=torch.zeros(8, 8))
