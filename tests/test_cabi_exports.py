"""CPU: the C-ABI library loads and exports every symbol include/links_b200.h declares (no compute calls)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "links_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(links_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    import __graft_entry__ as ge
    so = ge.build()
    from links_b200 import _cabi
    L = _cabi.load_library(so)
    syms = header_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(L, s), "missing export " + s
    # python binding table and header agree
    assert sorted(_cabi.ALL_SYMBOLS) == syms
    assert L.links_abi_version() == 1


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: constructing the engine without a CUDA device must raise."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from links_b200 import _cabi
    from links_b200.mlp import MlpSet
    with pytest.raises(_cabi.LinksError):
        MlpSet("lifter", [14], [{"downscale": 7, "angles": 1}], 8)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "links-3d-human-pose-estimation_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)


def header_prototypes():
    """name -> list of parameter declarations of every function prototype in include/links_b200.h."""
    src = open(os.path.join(ROOT, "include", "links_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t)\s+(links_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = [p.strip() for p in m.group(2).split(",") if p.strip() and p.strip() != "void"]
        out[m.group(1)] = params
    return out


def test_python_binding_matches_header_prototypes():
    """The ctypes table (links_b200/_cabi.py) declares every entry point with as many arguments as the header, with
    pointers where the header has pointers and scalars where it has scalars."""
    import ctypes as C
    from links_b200 import _cabi
    protos = header_prototypes()
    table = {n: list(a) + [C.c_void_p] for n, (_, a) in _cabi.SIGNATURES.items()}       # + stream
    table.update({n: list(a) for n, (_, a) in list(_cabi.PLAIN.items()) + list(_cabi.GEMM.items())})
    assert set(protos) == set(table)
    for name, params in protos.items():
        args = table[name]
        assert len(args) == len(params), (name, params, args)
        for decl, ct in zip(params, args):
            is_ptr_decl = "*" in decl
            is_ptr_ct = ct is C.c_void_p or hasattr(ct, "contents") or getattr(ct, "_type_", None) == "P" or \
                (isinstance(ct, type) and issubclass(ct, C._Pointer))
            assert is_ptr_decl == is_ptr_ct, (name, decl, ct)
            if not is_ptr_decl:
                if re.match(r"(const\s+)?float\b", decl):
                    assert ct is C.c_float, (name, decl, ct)
                elif re.match(r"(const\s+)?size_t\b", decl):
                    assert ct is C.c_size_t, (name, decl, ct)
                else:
                    assert ct in (C.c_int, C.c_uint), (name, decl, ct)
