"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports
every symbol include/densepoints_cuda.h declares; the product fails loudly without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os

import pytest

from densepoints_b200 import build as dpbuild
from densepoints_b200 import capi


@pytest.fixture(scope="module")
def built():
    return dpbuild.build_cuda()


def test_library_exports_every_declared_symbol(built):
    L = C.CDLL(built)
    names = capi.declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_abi_version_and_default_params(built):
    assert capi.lib().dp_abi_version() == 1
    p = capi.default_params()
    # the reference's constants (SURVEY section 5 "Config")
    assert p.score_threshold == 0.6 and p.minimum_visible_image == 3
    assert p.visible_threshold == 0.78 and p.candidate_threshold == 1.04
    assert p.grid_scale == 8 and p.max_patches_per_cell == 1
    assert list(p.nm_step) == [0.02, 0.2, 0.2] and p.nm_max_evals == 500 and p.nm_eps == 1e-4
    assert p.max_pops == 10_000_000


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.DpError):
        capi.Context(0)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "densepoints_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "dp_oracle" not in txt and "from oracle" not in txt and \
                    "import oracle" not in txt, f"{f} references the oracle"
