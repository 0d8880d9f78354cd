"""CPU: the C-ABI library loads and exports every symbol include/b200snark.h
declares (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import build as b200build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "b200snark.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    path = b200build.build_cuda()
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_fails_loudly_without_gpu():
    """No CPU fallback: on a box without a CUDA device every entry point errors."""
    import torch
    if torch.cuda.is_available():
        return
    from gosnark_b200 import _lib
    rc = _lib.lib().b200_init(0)
    assert rc == -1
    assert b"no CPU fallback" in _lib.lib().b200_last_error()


def test_go_binding_and_docs_name_only_declared_symbols():
    """Every C.b200_* the cgo sketch (go/) and INTEGRATION.md call is declared in include/b200snark.h."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "b200snark.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header)) | {"b200_pk_t", "b200_bases_t", "b200_r1cs_t"}
    used = set()
    for rel in ("INTEGRATION.md", os.path.join("go", "b200", "b200.go"), os.path.join("go", "groth16_generateproofs.go.txt")):
        path = os.path.join(root, rel)
        if os.path.exists(path):
            used |= set(re.findall(r"\bC\.(b200_[a-z0-9_]+)", open(path).read()))
    assert used, "no cgo calls found"
    assert used <= declared, sorted(used - declared)
