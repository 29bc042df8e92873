"""The specification of the opt-in extensions (oracle/flow2d_oracle_ext.c) checked on the CPU: with every extension
off it IS the reference restatement, and each extension behaves as numerical analysis says it must.  (SURVEY.md 8(f)
ranks 3-4; the reference has none of these features, so there is nothing upstream to compare with.)"""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def pair(synth):
    f0, f1, ut, vt = synth.make_pair(96, 80, 9, U1=1.0)
    lo = min(f0.min(), f1.min())
    return f0 - lo, f1 - lo, ut, vt  # non-negative intensities (the log term)


@pytest.mark.parametrize("constancy", [0, 1])
def test_all_extensions_off_is_the_reference_restatement(oracle, pair, constancy):
    f0, f1, _, _ = pair
    P = oracle.make_params(levels=8, outer=4, inner=5, alpha=20.0, sigma=1.0, median=3, constancy=constancy)
    u0, v0 = oracle.compute_flow(f0, f1, P)
    u1, v1, used = oracle.ext_compute_flow(f0, f1, P, oracle.make_ext())
    assert np.array_equal(u0, u1) and np.array_equal(v0, v1)
    assert used == [4] * len(used) and len(used) == len(oracle.level_table(96, 80, 0.9, 8))
    # omega = 1 and omega = 0 both mean "no relaxation factor"
    u2, v2, _ = oracle.ext_compute_flow(f0, f1, P, oracle.make_ext(omega=0.0))
    assert np.array_equal(u0, u2) and np.array_equal(v0, v2)


def _residual_after(oracle, f0, f1, ext, sweeps):
    z = np.zeros_like(f0)
    du, dv, phi, ksi, _ = oracle.ext_solve_level(f0, f1, z, z, 1.0, 1.0, oracle.make_params(outer=1, inner=sweeps, alpha=20.0), ext)
    return max(oracle.residual(f0, f1, z, z, du, dv, phi, ksi, 1.0, 1.0, 20.0))


def test_sor_converges_much_faster_than_jacobi(oracle, pair):
    """Red-black SOR with omega close to 2 removes the smooth error modes that Jacobi barely touches."""
    f0, f1 = pair[0][:40, :48].copy(), pair[1][:40, :48].copy()
    jac = _residual_after(oracle, f0, f1, oracle.make_ext(), 1500)
    gs = _residual_after(oracle, f0, f1, oracle.make_ext(scheme=oracle.RED_BLACK), 1500)
    sor = _residual_after(oracle, f0, f1, oracle.make_ext(scheme=oracle.RED_BLACK, omega=1.9), 1500)
    assert gs < jac and sor < 0.01 * jac, (jac, gs, sor)
    # damped Jacobi (omega < 1) is slower than plain Jacobi
    assert _residual_after(oracle, f0, f1, oracle.make_ext(omega=0.7), 300) > _residual_after(oracle, f0, f1, oracle.make_ext(), 300)


def test_jacobi_extension_path_agrees_with_the_reference_iteration(oracle, pair):
    """The tensor-plane form of the brightness term (extension solver, here forced by omega = 0.999999) differs from the
    reference's only in how ft^2 enters ksi (a separately rounded plane instead of an fma operand)."""
    f0, f1, _, _ = pair
    z = np.zeros_like(f0)
    P = oracle.make_params(outer=3, inner=5, alpha=20.0)
    du0, dv0, _, _ = oracle.solve_level(f0, f1, z, z, 1.0, 1.0, P)
    du1, dv1, _, _, _ = oracle.ext_solve_level(f0, f1, z, z, 1.0, 1.0, P, oracle.make_ext(omega=np.float32(0.9999999)))
    assert np.abs(du0 - du1).max() < 1e-3 and np.abs(dv0 - dv1).max() < 1e-3


def test_early_exit_counts(oracle, pair):
    f0, f1, _, _ = pair
    P = oracle.make_params(levels=10, scale=0.8, outer=12, inner=5, alpha=20.0, sigma=1.0, median=3)
    n = len(oracle.level_table(96, 80, 0.8, 10))
    _, _, none = oracle.ext_compute_flow(f0, f1, P, oracle.make_ext())
    _, _, huge = oracle.ext_compute_flow(f0, f1, P, oracle.make_ext(residual_tolerance=1e9))
    _, _, every3 = oracle.ext_compute_flow(f0, f1, P, oracle.make_ext(residual_tolerance=1e9, residual_check_every=3))
    u, v, some = oracle.ext_compute_flow(f0, f1, P, oracle.make_ext(residual_tolerance=1.0))
    assert none == [12] * n and huge == [1] * n and every3 == [3] * n
    assert all(1 <= k <= 12 for k in some) and min(some) < 12 and np.isfinite(u).all() and np.isfinite(v).all()


@pytest.mark.parametrize("term,gamma", [(1, 0.0), (2, 0.0), (3, 2.0)])
def test_tensor_terms(oracle, pair, term, gamma):
    f0, f1, ut, vt = pair
    J = oracle.ext_tensor(f0, f1, 1.0, 1.0, term, gamma)
    assert all(np.isfinite(j).all() for j in J)
    assert (J[0] >= 0).all() and (J[1] >= 0).all() and (J[5] >= 0).all()        # diagonal of a sum of outer products
    assert (J[0] * J[1] - J[2] * J[2] >= -1e-3 * (1 + J[0] * J[1])).all()      # leading 2x2 minor positive semi-definite
    if term == 3:
        B, G = oracle.ext_tensor(f0, f1, 1.0, 1.0, 0), oracle.ext_tensor(f0, f1, 1.0, 1.0, 1)
        for j, b, g in zip(J, B, G):
            assert np.allclose(j, b + gamma * g, rtol=1e-6, atol=1e-6)
    P = oracle.make_params(levels=8, outer=6, inner=5, alpha=10.0 if term != 2 else 0.05, sigma=1.0, median=3)
    u, v, _ = oracle.ext_compute_flow(f0, f1, P, oracle.make_ext(data_term=term, gamma=gamma))
    assert np.isfinite(u).all() and np.isfinite(v).all()
    if term == 3:  # the combined term must still find the synthetic motion
        assert np.hypot(u - ut, v - vt).mean() < 0.5 * np.hypot(ut, vt).mean()


def test_cascaded_restriction_is_close_but_not_identical(oracle, pair):
    f0, f1, _, _ = pair
    P = oracle.make_params(levels=10, outer=5, inner=5, alpha=20.0, sigma=1.0, median=3)
    u0, v0 = oracle.compute_flow(f0, f1, P)
    u1, v1, _ = oracle.ext_compute_flow(f0, f1, P, oracle.make_ext(cascaded_restriction=1))
    d = np.hypot(u0 - u1, v0 - v1)
    assert d.max() > 0 and d.mean() < 0.25
    # one level only: nothing to restrict, identical
    P1 = oracle.make_params(levels=1, outer=5, inner=5, alpha=20.0, sigma=1.0, median=3)
    a = oracle.compute_flow(f0, f1, P1)
    b = oracle.ext_compute_flow(f0, f1, P1, oracle.make_ext(cascaded_restriction=1))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_extension_fields_default_to_zero(pkg):
    p = pkg.default_params()
    assert (p.scheme, p.omega, p.data_term, p.gamma, p.residual_tolerance, p.residual_check_every, p.cascaded_restriction) == (0, 0, 0, 0, 0, 0, 0)
