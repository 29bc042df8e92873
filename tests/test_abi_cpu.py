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


def test_live_handle_count_is_exported(pkg):
    """flow2d_live_handles (the scheduler's "several pairs in flight?" signal) through the package"""
    assert pkg.live_handles(0) >= 0 and pkg.live_handles(-1) == 0 and pkg.live_handles(1000) == 0


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


def test_level_table_random_sizes(pkg, oracle):
    """Randomised: flow2d_max_warp_level / flow2d_level_geometry (host-only fp32 arithmetic of
    optical_flow_base_2d.cpp:36-59, optical_flow_2d.cpp:268-272) agree with the oracle for any size and scale."""
    rng = np.random.default_rng(42)
    for _ in range(300):
        W, H = int(rng.integers(4, 9000)), int(rng.integers(4, 9000))
        sf = float(np.float32(rng.uniform(0.3, 0.97)))
        assert pkg.max_warp_level(W, H, sf) == oracle.max_warp_level(W, H, sf), (W, H, sf)
        a, b = pkg.level_table(W, H, sf, 1000), oracle.level_table(W, H, sf, 1000)
        assert len(a) == len(b) and all(x[:2] == y[:2] and x[2].tobytes() == y[2].tobytes() and x[3].tobytes() == y[3].tobytes()
                                        for x, y in zip(a, b)), (W, H, sf)


def test_null_handle_is_an_error_not_a_crash(pkg):
    """Every entry point that takes a handle answers FLOW2D_ERR_INVALID_ARGUMENT (-1) for NULL (the reference
    dereferences uninitialised operation pointers in that situation)."""
    import ctypes as C
    L = pkg.binding.lib() if hasattr(pkg, "binding") else None
    if L is None:
        from cuda_flow2d_b200 import binding
        L = binding.lib()
    null = C.c_void_p(0)
    p = pkg.default_params()
    assert L.flow2d_compute(null, null, null, null, null, C.byref(p)) == -1
    assert L.flow2d_compute_device(null, null, null, null, null, C.byref(p)) == -1
    assert L.flow2d_synchronize(null) == -1
    assert L.flow2d_set_stream(null, null) == -1
    assert L.flow2d_last_launch_counts(null, None) == -1
    assert L.flow2d_pitch_elems(null) == 0 and L.flow2d_width(null) == 0
    assert L.flow2d_destroy(null) == 0  # destroying nothing is fine
