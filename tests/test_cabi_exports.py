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
