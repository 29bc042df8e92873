"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/flow2d.h declares, its host-only level table is bit exact with the oracle, and it refuses
to compute without a device (no CPU fallback exists)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "flow2d.h")).read()
    return sorted(set(re.findall(r"FLOW2D_API[^;(]*?\b(flow2d_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    names = _declared_symbols()
    assert len(names) >= 20
    L = ctypes.CDLL(pkg.lib_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_defaults(pkg):
    assert "sm_100a" in pkg.version()
    p = pkg.default_params()
    # src/main.cpp:70-80
    assert (p.warp_levels_count, p.outer_iterations_count, p.inner_iterations_count, p.median_radius) == (50, 40, 5, 5)
    assert (p.warp_scale_factor, p.equation_alpha, p.gaussian_sigma) == (np.float32(0.9), 35.0, 1.5)
    assert p.equation_smoothness == np.float32(0.001) and p.equation_data == np.float32(0.001)


@pytest.mark.parametrize("W,H", [(584, 388), (128, 128), (1024, 1024), (2048, 2048), (8192, 8192), (640, 480), (37, 1001), (5, 4)])
@pytest.mark.parametrize("sf", [0.9, 0.5, 0.95, 0.75])
def test_level_table_bit_exact_with_oracle(pkg, oracle, W, H, sf):
    assert pkg.max_warp_level(W, H, sf) == oracle.max_warp_level(W, H, sf)
    a = pkg.level_table(W, H, sf, 100)
    b = oracle.level_table(W, H, sf, 100)
    assert len(a) == len(b)
    for (cw, ch, hx, hy), (ow, oh, ohx, ohy) in zip(a, b):
        assert (cw, ch) == (ow, oh)
        assert hx.tobytes() == ohx.tobytes() and hy.tobytes() == ohy.tobytes()


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the handle cannot be created: the product has no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.Flow2DError) as e:
        pkg.Flow2D(64, 64)
    assert e.value.code == -3  # FLOW2D_ERR_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "cuda-flow2d_b200")):
        if os.sep + "build" in base or os.sep + "lib" in base or os.sep + "bin" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                t = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"liboracle|flow2d_oracle|from oracle|import oracle|oracle/", t):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
