"""The C-ABI library loads and exports every symbol include/pinn_b200.h declares; descriptor layout and
support table are consistent between Python and C (no compute calls -- runs without a GPU)."""
import ctypes
import os
import re

import pytest

from pinnacle_b200 import capi, kinds, problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.LIB_PATH):
        import subprocess
        import sys
        subprocess.run([sys.executable, os.path.join(ROOT, "pinnacle_b200", "csrc", "build.py")], check=True)
    return capi.load()


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "pinn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pinn_[a-z_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    names = _header_functions()
    assert set(names) == set(capi.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_descriptor_layout_matches_header():
    assert ctypes.sizeof(kinds.CKindDesc) == 4 * 7 + 4 * kinds.PINN_MAX_DIM + 4 * kinds.PINN_MAX_AUXSEL + 4 * kinds.PINN_MAX_CONSTS
    txt = open(os.path.join(ROOT, "include", "pinn_b200.h")).read()
    for name in ("PINN_MAX_DIM", "PINN_MAX_CONSTS", "PINN_MAX_AUXSEL", "PINN_MAX_PDE"):
        assert int(re.search(rf"#define {name} (\d+)", txt).group(1)) == getattr(kinds, name)
    for m in re.finditer(r"PINN_KIND_([A-Z0-9_]+) = (\d+)", txt):
        assert getattr(kinds, "KIND_" + m.group(1)) == int(m.group(2))


def test_param_count_and_support_table(lib):
    import numpy as np
    for name, builder in problems.BUILDERS.items():
        if name in ("Poisson2D_ManyArea",):
            spec = builder(np.ones((5, 5)), np.ones((5, 5, 2, 2)))
        elif name in ("Heat2D_VaryingCoef", "Wave2D_Heterogeneous"):
            spec = builder(np.array([[0.0, 0.0, 1.0], [1.0, 1.0, 2.0], [0.0, 1.0, 1.5], [1.0, 0.0, 1.2]]))
        else:
            spec = builder()
        c = spec.desc.to_c()
        assert lib.pinn_param_count(ctypes.byref(c)) == spec.desc.n_params
        assert capi.supported(spec.desc), (name, capi.last_error())
    bad = kinds.KindDesc(kinds.KIND_BURGERS1D, 2, 1, (1, 1))
    assert not capi.supported(bad)
    assert "no kernel" in capi.last_error()


def test_build_info(lib):
    assert "sm_100a" in capi.build_info()
