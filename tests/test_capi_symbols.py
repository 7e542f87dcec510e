"""The C-ABI library loads and exports every symbol include/gears_physics_cuda.h declares (no GPU needed)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "gears_physics_cuda.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(gp_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from gears_b200 import capi
    L = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in the header but not exported"
    assert sorted(capi.SYMBOLS) == syms
    assert L.gp_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gears_b200.world import GpError, PhysicsWorld
    with pytest.raises(GpError) as e:
        PhysicsWorld()
    assert e.value.status == -3


def test_product_code_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "gears_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in txt and "gears_oracle" not in txt and "from oracle" not in txt, f
