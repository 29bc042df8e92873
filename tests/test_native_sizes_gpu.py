"""Parity at the NATIVE sizes of BASELINE.json's configs, against the reference's own CUDA build
(oracle/_ref/ref_harness, OpticalFlow2D::ComputeFlow of src/optical_flow/optical_flow_2d.cpp:142-569)
on the same frames, same settings, same box.

  C2  1024x1024, one level, 500 Jacobi sweeps, automatic schedule           array_equal
  C4  one 1024x1024 radiography-like pair, 50 levels, 40x5 (bench.py's frames) array_equal
  C3  2048x2048 full pyramid, Grey                                           array_equal
  C3  2048x2048, GRADIENT constancy, scale-0.5 pyramid (all levels multiples   array_equal
      of the reference's 16x8 block: no undefined cells upstream)
  C3  2048x2048 full 0.9 pyramid, GRADIENT constancy                         kernel == oracle bit for bit; the reference
                                                                             itself reads uninitialised shared memory here
                                                                             (SURVEY.md F5): distance measured, loosely bounded
  C5  8192x8192 full pyramid, Grey, once                                     array_equal

These are multi-wave launches of every solve kernel (704-CTA grids, PDL between passes, CUDA graph replay),
i.e. the paths bench.py times.  Grey mode is bit exact with the reference by construction.
"""
import numpy as np
import pytest

from conftest import epd_stats

pytestmark = pytest.mark.gpu

C2 = dict(levels=1, scale=0.5, outer=1, inner=500, alpha=0.25, e_smooth=1.0, e_data=1000.0, median=1, sigma=0.0)
FULL = dict(levels=50, scale=0.9, outer=40, inner=5, alpha=35.0, e_smooth=0.001, e_data=0.001, median=5, sigma=1.5)
C3 = dict(FULL, alpha=3.5)


@pytest.fixture(scope="module")
def torch_():
    import torch
    assert torch.cuda.is_available()
    return torch


def _ours(pkg, f0, f1, cfg, constancy=0, twice=True):
    h, w = f0.shape
    fl = pkg.Flow2D(w, h, constancy=constancy)
    p = pkg.default_params(**cfg)
    u, v = fl.compute(f0, f1, p)
    if twice:  # the replayed CUDA graph must give the same bits as the first (capturing) call
        u2, v2 = fl.compute(f0, f1, p)
        assert np.array_equal(u, u2) and np.array_equal(v, v2), "graph replay differs from the first call"
    st, counts = fl.stats(), fl.launch_counts()
    fl.destroy()
    return u, v, st, counts


def test_c2_native_vs_reference(pkg, synth, ref, torch_):
    """The configuration bench.py's `extra.c2` times: 1024x1024, single level, 500 sweeps, S chosen by the scheduler."""
    f0, f1, _, _ = synth.make_pair(1024, 1024, 1001, U0=(0.3, -0.2), U1=0.5, L=256.0)
    ru, rv, _ = ref.flow(f0, f1, C2)
    u, v, st, counts = _ours(pkg, f0, f1, C2)
    s = epd_stats(u, v, ru, rv)
    print("C2 native vs reference:", s, st, counts)
    assert counts.get("solve_pass", 0) > 50  # the temporally blocked 64x48 kernel, multi-wave grid
    assert s["exact"], s


def test_c4_pair_native_vs_reference(pkg, synth, ref, torch_):
    """One pair of bench.py's default workload (C4): 1024x1024, sigma_n = 2 noise, contrast 0.3, 50 levels, 40x5."""
    f0, f1, _, _ = synth.make_pair(1024, 1024, 4000, U0=(0.0, 0.0), U1=4.0, L=384.0, contrast=0.3, noise=2.0)
    ru, rv, _ = ref.flow(f0, f1, FULL)
    u, v, st, counts = _ours(pkg, f0, f1, FULL)
    s = epd_stats(u, v, ru, rv)
    print("C4 pair native vs reference:", s, st, counts)
    assert st["levels_run"] == 50
    assert s["exact"], s


def test_c3_grey_native_vs_reference(pkg, synth, ref, torch_):
    f0, f1, _, _ = synth.make_pair(2048, 2048, 2001, U0=(3.0, -2.0), U1=6.0, L=512.0)
    ru, rv, _ = ref.flow(f0, f1, C3)
    u, v, st, counts = _ours(pkg, f0, f1, C3, twice=False)
    s = epd_stats(u, v, ru, rv)
    print("C3-Grey native vs reference:", s, st, counts)
    assert s["exact"], s


def test_c3_gradient_native_vs_reference_clean_pyramid(pkg, synth, ref, torch_):
    """Gradient constancy at 2048x2048 where the reference IS a function of its inputs: a scale-0.5 pyramid
    2048, 1024, ..., 16 -- every level a multiple of its 16x8 CUDA block, so solve_2d_grad never reads the
    uninitialised shared-memory cells of SURVEY.md F5 -- through the whole pipeline (blur, restriction, prolongation,
    warp, 40x5 robust iterations with the tile-coupled gradient tensor, median).  Bit for bit."""
    f0, f1, _, _ = synth.make_pair(2048, 2048, 2001, U0=(3.0, -2.0), U1=6.0, L=512.0)
    cfg = dict(C3, levels=8, scale=0.5)
    ru, rv, _ = ref.flow(f0, f1, cfg, constancy=1)
    u, v, st, counts = _ours(pkg, f0, f1, cfg, constancy=1, twice=False)
    s = epd_stats(u, v, ru, rv)
    print("C3-gradient, clean pyramid (8 levels, scale 0.5) vs reference:", s, st, counts)
    assert st["levels_run"] == 8 and counts.get("grad_tensor", 0) == 8
    assert s["exact"], s


def test_c3_gradient_native_vs_reference(pkg, oracle, synth, ref, torch_):
    """BASELINE.json configs[2] with the 0.9 pyramid.  The reference's solve_2d_grad reads uninitialised shared memory
    next to partial 16x8 blocks (last column / row of every level whose size is not a multiple of 16x8,
    i.e. all 49 coarse levels), so its result is not a function of its inputs: whatever a previous block left in shared
    memory enters the tensor at the right / bottom border of every coarse level and is prolongated over the frame.
    Our definition of those cells (own value) is pinned by the oracle bit for bit; the distance to the reference's
    run is measured and bounded loosely (measured on B200: mean 1.6e-2 px, i.e. 16x the north-star mean that the
    clean pyramid above and every Grey configuration meet with 0.0)."""
    f0, f1, ut, vt = synth.make_pair(2048, 2048, 2001, U0=(3.0, -2.0), U1=6.0, L=512.0)
    u, v, st, counts = _ours(pkg, f0, f1, C3, constancy=1, twice=False)
    ou, ov = oracle.compute_flow(f0, f1, oracle.make_params(constancy=oracle.GRADIENT, **C3))
    so = epd_stats(u, v, ou, ov)
    print("C3-gradient native kernel vs oracle:", so, st, counts)
    assert so["exact"], so
    ru, rv, _ = ref.flow(f0, f1, C3, constancy=1)
    s = epd_stats(u, v, ru, rv)
    print("C3-gradient native vs reference, full frame:", s)
    print("C3-gradient: mean endpoint error against the synthetic ground truth: ours %.4f px, reference %.4f px" %
          (float(np.hypot(u - ut, v - vt).mean()), float(np.hypot(ru - ut, rv - vt).mean())))
    assert s["mean"] <= 5e-2, s


def test_c5_grey_native_vs_reference(pkg, synth, ref, torch_):
    """BASELINE.json configs[4] on one GPU: 8192x8192, full pyramid (76 possible levels, 50 used)."""
    w = h = 8192
    g0, g1 = synth.make_pair_torch(w, h, 5001, "cuda:0", U0=(0.0, 0.0), U1=8.0, L=2048.0)
    f0, f1 = g0.cpu().numpy(), g1.cpu().numpy()
    del g0, g1
    torch_.cuda.empty_cache()
    u, v, st, counts = _ours(pkg, f0, f1, C3, twice=False)
    torch_.cuda.empty_cache()
    ru, rv, _ = ref.flow(f0, f1, C3)
    s = epd_stats(u, v, ru, rv)
    print("C5-Grey native vs reference:", s, st, counts)
    assert s["exact"], s
