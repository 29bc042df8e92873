"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this
module.  The product package (cuda-flow2d_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GREY, GRADIENT = 0, 1


class Params(C.Structure):
    """Mirror of `oracle_params` (flow2d_oracle.h); field meaning = main.cpp:70-80 of the reference."""
    _fields_ = [
        ("warp_levels_count", C.c_size_t),
        ("warp_scale_factor", C.c_float),
        ("outer_iterations_count", C.c_size_t),
        ("inner_iterations_count", C.c_size_t),
        ("equation_alpha", C.c_float),
        ("equation_smoothness", C.c_float),
        ("equation_data", C.c_float),
        ("median_radius", C.c_size_t),
        ("gaussian_sigma", C.c_float),
        ("constancy", C.c_int),
    ]


def make_params(levels=50, scale=0.9, outer=40, inner=5, alpha=35.0, e_smooth=0.001, e_data=0.001,
                median=5, sigma=1.5, constancy=GREY):
    return Params(levels, scale, outer, inner, alpha, e_smooth, e_data, median, sigma, constancy)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("flow2d_oracle.c", "flow2d_oracle_ext.c", "flow2d_oracle.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        sz = C.c_size_t
        L.oracle_max_warp_level.restype = sz
        L.oracle_max_warp_level.argtypes = [sz, sz, C.c_float]
        L.oracle_level_geometry.argtypes = [sz, sz, C.c_float, C.c_int, C.POINTER(sz), C.POINTER(sz), fp, fp]
        L.oracle_gauss_taps.restype = C.c_int
        L.oracle_gauss_taps.argtypes = [C.c_float, fp]
        L.oracle_blur.argtypes = [fp, fp, sz, sz, sz, C.c_float]
        L.oracle_resample.argtypes = [fp, sz, sz, fp, sz, sz, sz]
        L.oracle_resample_cells.argtypes = [sz, sz, sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_warp.argtypes = [fp, fp, fp, fp, sz, sz, sz, C.c_float, C.c_float, fp]
        L.oracle_phi_ksi.argtypes = [fp] * 6 + [sz, sz, sz] + [C.c_float] * 4 + [fp, fp]
        L.oracle_sweep_grey.argtypes = [fp] * 8 + [sz, sz, sz] + [C.c_float] * 3 + [fp, fp]
        L.oracle_sweep_grad.argtypes = [fp] * 8 + [sz, sz, sz] + [C.c_float] * 3 + [fp, fp]
        L.oracle_solve_level.argtypes = [fp] * 10 + [sz, sz, sz, C.c_float, C.c_float, C.POINTER(Params)]
        L.oracle_add.argtypes = [fp, fp, sz, sz, sz]
        L.oracle_median.restype = C.c_int
        L.oracle_median.argtypes = [fp, fp, sz, sz, sz, sz]
        L.oracle_compute_flow.restype = C.c_int
        L.oracle_compute_flow.argtypes = [fp, fp, sz, sz, C.POINTER(Params), fp, fp]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def max_warp_level(w, h, sf):
    return int(lib().oracle_max_warp_level(w, h, sf))


def level_geometry(W, H, sf, level):
    cw, ch, hx, hy = C.c_size_t(), C.c_size_t(), C.c_float(), C.c_float()
    lib().oracle_level_geometry(W, H, sf, level, C.byref(cw), C.byref(ch), C.byref(hx), C.byref(hy))
    return cw.value, ch.value, np.float32(hx.value), np.float32(hy.value)


def level_table(W, H, sf, levels):
    """[(cw, ch, hx, hy)] from the coarsest used level down to level 0 (optical_flow_2d.cpp:188-189,267-272)."""
    n = min(levels, max_warp_level(W, H, sf))
    return [level_geometry(W, H, sf, l) for l in range(n - 1, -1, -1)]


def gauss_taps(sigma):
    t = np.zeros(64, np.float32)
    r = lib().oracle_gauss_taps(sigma, _p(t))
    return t[: 2 * r + 1].copy(), r


def blur(img, sigma):
    img = _f32(img)
    h, w = img.shape
    out = np.empty_like(img)
    lib().oracle_blur(_p(img), _p(out), w, h, w, sigma)
    return out


def resample(img, ow, oh):
    """Area resampling of a dense (ih, iw) image to (oh, ow)."""
    img = _f32(img)
    ih, iw = img.shape
    pitch = max(iw, ow)
    src = np.zeros((max(ih, oh), pitch), np.float32)
    src[:ih, :iw] = img
    dst = np.zeros_like(src)
    lib().oracle_resample(_p(src), iw, ih, _p(dst), ow, oh, pitch)
    return dst[:oh, :ow].copy()


def resample_cells(in_n, out_n):
    li, ri = C.c_int(), C.c_int()
    out = np.zeros((out_n, 2), np.int32)
    for x in range(out_n):
        lib().oracle_resample_cells(in_n, out_n, x, C.byref(li), C.byref(ri))
        out[x] = (li.value, ri.value)
    return out


def warp(f0, f1, u, v, hx, hy):
    f0, f1, u, v = map(_f32, (f0, f1, u, v))
    h, w = f0.shape
    out = np.empty_like(f0)
    lib().oracle_warp(_p(f0), _p(f1), _p(u), _p(v), w, h, w, hx, hy, _p(out))
    return out


def phi_ksi(f0, f1, u, v, du, dv, hx, hy, e_smooth, e_data):
    f0, f1, u, v, du, dv = map(_f32, (f0, f1, u, v, du, dv))
    h, w = f0.shape
    phi, ksi = np.empty_like(f0), np.empty_like(f0)
    lib().oracle_phi_ksi(_p(f0), _p(f1), _p(u), _p(v), _p(du), _p(dv), w, h, w, hx, hy, e_smooth, e_data, _p(phi), _p(ksi))
    return phi, ksi


def sweep(f0, f1, u, v, du, dv, phi, ksi, hx, hy, alpha, constancy=GREY):
    f0, f1, u, v, du, dv, phi, ksi = map(_f32, (f0, f1, u, v, du, dv, phi, ksi))
    h, w = f0.shape
    odu, odv = np.empty_like(f0), np.empty_like(f0)
    fn = lib().oracle_sweep_grad if constancy == GRADIENT else lib().oracle_sweep_grey
    fn(_p(f0), _p(f1), _p(u), _p(v), _p(du), _p(dv), _p(phi), _p(ksi), w, h, w, hx, hy, alpha, _p(odu), _p(odv))
    return odu, odv


def solve_level(f0, f1, u, v, hx, hy, params):
    """cuda_operation_solve_2d.cpp:106-315 on dense arrays; returns (du, dv, phi, ksi)."""
    f0, f1, u, v = map(_f32, (f0, f1, u, v))
    h, w = f0.shape
    du, dv, phi, ksi, t0, t1 = (np.zeros_like(f0) for _ in range(6))
    lib().oracle_solve_level(_p(f0), _p(f1), _p(u), _p(v), _p(du), _p(dv), _p(phi), _p(ksi), _p(t0), _p(t1),
                             w, h, w, hx, hy, C.byref(params))
    return du, dv, phi, ksi


def residual(f0, f1, u, v, du, dv, phi, ksi, hx, hy, alpha, constancy=GREY):
    """EXTENSION (no reference counterpart): RMS residual of the lagged linear system; returns (rms_u, rms_v)."""
    f0, f1, u, v, du, dv, phi, ksi = map(_f32, (f0, f1, u, v, du, dv, phi, ksi))
    h, w = f0.shape
    ru, rv = C.c_double(), C.c_double()
    fn = lib().oracle_residual
    fn.restype = None
    fn.argtypes = [C.c_void_p] * 8 + [C.c_size_t, C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_int,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
    fn(_p(f0), _p(f1), _p(u), _p(v), _p(du), _p(dv), _p(phi), _p(ksi), w, h, w, hx, hy, alpha, constancy, C.byref(ru), C.byref(rv))
    return ru.value, rv.value


def median(img, radius):
    img = _f32(img)
    h, w = img.shape
    out = np.zeros_like(img)
    rc = lib().oracle_median(_p(img), _p(out), w, h, w, radius)
    return out, rc


def compute_flow(f0, f1, params):
    f0, f1 = _f32(f0), _f32(f1)
    h, w = f0.shape
    u, v = np.empty_like(f0), np.empty_like(f0)
    rc = lib().oracle_compute_flow(_p(f0), _p(f1), w, h, C.byref(params), _p(u), _p(v))
    if rc != 0:
        raise RuntimeError("oracle_compute_flow failed: %d" % rc)
    return u, v


# ---- EXTENSIONS beyond the reference (flow2d_oracle_ext.c): the specification the opt-in features are checked against ----
JACOBI, RED_BLACK = 0, 1
TERM_DEFAULT, TERM_GRADIENT, TERM_LOG_GRADIENT, TERM_COMBINED = 0, 1, 2, 3


class Ext(C.Structure):
    """Mirror of `oracle_ext` (flow2d_oracle.h)."""
    _fields_ = [
        ("scheme", C.c_int),
        ("omega", C.c_float),
        ("data_term", C.c_int),
        ("gamma", C.c_float),
        ("residual_tolerance", C.c_float),
        ("residual_check_every", C.c_int),
        ("cascaded_restriction", C.c_int),
    ]


def make_ext(scheme=JACOBI, omega=1.0, data_term=TERM_DEFAULT, gamma=0.0, residual_tolerance=0.0, residual_check_every=1,
             cascaded_restriction=0):
    return Ext(scheme, omega, data_term, gamma, residual_tolerance, residual_check_every, cascaded_restriction)


def ext_tensor(f0, f1w, hx, hy, data_term, gamma=0.0):
    """Six tensor planes J11 J22 J12 J13 J23 J33 of a level."""
    f0, f1w = _f32(f0), _f32(f1w)
    h, w = f0.shape
    J = [np.zeros_like(f0) for _ in range(6)]
    arr = (C.POINTER(C.c_float) * 6)(*[_p(j) for j in J])
    fn = lib().oracle_ext_tensor
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]
    fn(_p(f0), _p(f1w), w, h, w, hx, hy, data_term, gamma, arr)
    return J


def ext_solve_level(f0, f1w, u, v, hx, hy, params, ext):
    """One level with the extensions; returns (du, dv, phi, ksi, outer iterations that ran)."""
    f0, f1w, u, v = map(_f32, (f0, f1w, u, v))
    h, w = f0.shape
    du, dv, phi, ksi = (np.zeros_like(f0) for _ in range(4))
    fn = lib().oracle_ext_solve_level
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 8 + [C.c_size_t] * 3 + [C.c_float, C.c_float, C.POINTER(Params), C.POINTER(Ext)]
    used = fn(_p(f0), _p(f1w), _p(u), _p(v), _p(du), _p(dv), _p(phi), _p(ksi), w, h, w, hx, hy, C.byref(params), C.byref(ext))
    return du, dv, phi, ksi, int(used)


def ext_compute_flow(f0, f1, params, ext):
    """The whole path with the extensions; returns (u, v, [outer iterations per level, coarsest first])."""
    f0, f1 = _f32(f0), _f32(f1)
    h, w = f0.shape
    u, v = np.empty_like(f0), np.empty_like(f0)
    used = (C.c_int * 256)()
    fn = lib().oracle_ext_compute_flow
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(Params), C.POINTER(Ext), C.c_void_p, C.c_void_p,
                   C.c_void_p, C.c_int]
    n = fn(_p(f0), _p(f1), w, h, C.byref(params), C.byref(ext), _p(u), _p(v), used, 256)
    if n < 0:
        raise RuntimeError("oracle_ext_compute_flow failed: %d" % n)
    return u, v, [int(used[i]) for i in range(min(n, 256))]


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))
