"""ctypes binding of include/flow2d.h.  Plumbing only: torch supplies device memory and streams."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GREY, GRADIENT = 0, 1
KERNEL_KINDS = 14  # FLOW2D_KERNEL_KINDS
JACOBI, RED_BLACK = 0, 1  # FLOW2D_SCHEME_*
TERM_DEFAULT, TERM_GRADIENT, TERM_LOG_GRADIENT, TERM_COMBINED = 0, 1, 2, 3  # FLOW2D_TERM_*
MAX_LEVELS = 256   # FLOW2D_MAX_LEVELS

OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_NO_DEVICE, ERR_OUT_OF_MEMORY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5


class Flow2DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("flow2d error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    """`flow2d_params` (include/flow2d.h): the nine ComputeFlow parameters + scheduling knobs."""
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
        ("sweeps_per_pass", C.c_int),
        ("resident_levels", C.c_int),
        ("throughput_mode", C.c_int),
        ("report_residuals", C.c_int),
        # opt-in extensions beyond the reference (all zero = reference behaviour)
        ("scheme", C.c_int),
        ("omega", C.c_float),
        ("data_term", C.c_int),
        ("gamma", C.c_float),
        ("residual_tolerance", C.c_float),
        ("residual_check_every", C.c_int),
        ("cascaded_restriction", C.c_int),
        ("report_level_times", C.c_int),
    ]


def lib_path():
    return os.path.join(_HERE, "lib", "libflow2d_b200.so")


def build(force=False):
    """Compile the sm_100a library (and the CLI) in-tree with the package Makefile."""
    cmd = ["make", "-C", _HERE, "-j4"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return lib_path()


def lib():
    """Load the C-ABI library.  Fails loudly when it has not been built: there is no fallback."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise Flow2DError(ERR_NO_DEVICE, "%s is missing -- run __graft_entry__.build() (make -C cuda-flow2d_b200)" % path)
        L = C.CDLL(path)
        fp, sz, vp = C.POINTER(C.c_float), C.c_size_t, C.c_void_p
        L.flow2d_version.restype = C.c_char_p
        L.flow2d_default_params.argtypes = [C.POINTER(Params)]
        L.flow2d_create.argtypes = [C.POINTER(vp), C.c_int, sz, sz, C.c_int]
        L.flow2d_destroy.argtypes = [vp]
        L.flow2d_last_error.restype = C.c_char_p
        L.flow2d_last_error.argtypes = [vp]
        for name in ("flow2d_pitch_elems", "flow2d_width", "flow2d_height"):
            getattr(L, name).restype = sz
            getattr(L, name).argtypes = [vp]
        L.flow2d_set_stream.argtypes = [vp, vp]
        L.flow2d_get_stream.restype = vp
        L.flow2d_get_stream.argtypes = [vp]
        L.flow2d_compute.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Params)]
        L.flow2d_compute_device.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Params)]
        L.flow2d_compute_async.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Params)]
        L.flow2d_synchronize.argtypes = [vp]
        L.flow2d_prepare.argtypes = [vp, C.POINTER(Params)]
        L.flow2d_last_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_int), fp]
        L.flow2d_last_launch_counts.argtypes = [vp, C.POINTER(C.c_longlong)]
        L.flow2d_graph_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        dp = C.POINTER(C.c_double)
        L.flow2d_level_residuals.argtypes = [vp, dp, dp, C.c_int, C.POINTER(C.c_int)]
        L.flow2d_level_times.argtypes = [vp, fp, fp, C.c_int, C.POINTER(C.c_int)]
        L.flow2d_level_outer_iterations.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
        L.flow2d_stage_residual.argtypes = [vp] + [vp] * 8 + [C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.POINTER(Params), dp, dp]
        L.flow2d_kernel_kind_name.argtypes = [C.c_int]
        L.flow2d_kernel_kind_name.restype = C.c_char_p
        L.flow2d_max_warp_level.restype = sz
        L.flow2d_max_warp_level.argtypes = [sz, sz, C.c_float]
        L.flow2d_level_geometry.argtypes = [sz, sz, C.c_float, C.c_int, C.POINTER(sz), C.POINTER(sz), fp, fp]
        L.flow2d_stage_blur.argtypes = [vp, vp, vp, sz, sz, C.c_float]
        L.flow2d_stage_resample.argtypes = [vp, vp, sz, sz, vp, sz, sz]
        L.flow2d_stage_warp.argtypes = [vp, vp, vp, vp, vp, vp, sz, sz, C.c_float, C.c_float]
        L.flow2d_stage_solve.argtypes = [vp] * 9 + [sz, sz, C.c_float, C.c_float, C.POINTER(Params)]
        L.flow2d_stage_add.argtypes = [vp, vp, vp, sz, sz]
        L.flow2d_stage_median.argtypes = [vp, vp, vp, sz, sz, sz]
        L.flow2d_stage_add_median.argtypes = [vp, vp, vp, vp, sz, sz, sz]
        L.flow2d_debug_timing.argtypes = [vp, vp]
        L.flow2d_compute_slab_device.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Params)]
        L.flow2d_stage_solve_slab.argtypes = [vp] * 7 + [sz, sz, C.c_float, C.c_float, C.POINTER(Params), C.POINTER(C.c_int)]
        L.flow2d_slab_mailbox.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
        L.flow2d_slab_export.argtypes = [vp, C.c_char_p]
        L.flow2d_slab_import.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        L.flow2d_slab_connect.argtypes = [vp, C.c_int, C.c_int, vp, vp, sz]
        L.flow2d_slab_rows.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz)]
        L.flow2d_slab_status.argtypes = [vp]
        L.flow2d_slab_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


def version():
    return lib().flow2d_version().decode()


def default_params(**over):
    """Reference argv-form defaults (src/main.cpp:70-80), optionally overridden by short names."""
    p = Params()
    lib().flow2d_default_params(C.byref(p))
    short = {"levels": "warp_levels_count", "scale": "warp_scale_factor", "outer": "outer_iterations_count",
             "inner": "inner_iterations_count", "alpha": "equation_alpha", "e_smooth": "equation_smoothness",
             "e_data": "equation_data", "median": "median_radius", "sigma": "gaussian_sigma"}
    for k, v in over.items():
        setattr(p, short.get(k, k), v)
    return p


def live_handles(device=0):
    """Handles of this process alive on `device` (flow2d_live_handles): with FLOW2D_CLUSTER unset, 4 or more switch the
    mid-size levels to the thread-block-cluster solve."""
    return int(lib().flow2d_live_handles(device))


def max_warp_level(w, h, sf):
    return int(lib().flow2d_max_warp_level(w, h, sf))


def level_geometry(W, H, sf, level):
    cw, ch, hx, hy = C.c_size_t(), C.c_size_t(), C.c_float(), C.c_float()
    rc = lib().flow2d_level_geometry(W, H, sf, level, C.byref(cw), C.byref(ch), C.byref(hx), C.byref(hy))
    if rc != OK:
        raise Flow2DError(rc, "flow2d_level_geometry")
    return cw.value, ch.value, np.float32(hx.value), np.float32(hy.value)


def level_table(W, H, sf, levels):
    n = min(levels, max_warp_level(W, H, sf))
    return [level_geometry(W, H, sf, l) for l in range(n - 1, -1, -1)]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Flow2D:
    """One `flow2d_handle`: the replacement of an initialised reference `OpticalFlow2D` object."""

    def __init__(self, width, height, constancy=GREY, device=0):
        self._h = C.c_void_p(0)
        self.width, self.height, self.constancy, self.device = int(width), int(height), constancy, int(device)
        rc = lib().flow2d_create(C.byref(self._h), self.device, self.width, self.height, constancy)
        if rc != OK:
            self._h = C.c_void_p(0)
            raise Flow2DError(rc, "flow2d_create(%dx%d, device %d) failed" % (width, height, device))
        self.pitch = int(lib().flow2d_pitch_elems(self._h))

    # -- lifetime --
    def destroy(self):
        if self._h:
            lib().flow2d_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise Flow2DError(rc, lib().flow2d_last_error(self._h).decode())

    def set_stream(self, cuda_stream_handle):
        self._check(lib().flow2d_set_stream(self._h, C.c_void_p(cuda_stream_handle or 0)))

    def stats(self):
        n, lv, ms = C.c_longlong(), C.c_int(), C.c_float()
        self._check(lib().flow2d_last_stats(self._h, C.byref(n), C.byref(lv), C.byref(ms)))
        return {"kernel_launches": n.value, "levels_run": lv.value, "device_ms": ms.value}

    def graph_stats(self):
        """(captures, replays) of the handle's CUDA-graph cache since it was created (flow2d_graph_stats)."""
        a, b = C.c_longlong(), C.c_longlong()
        self._check(lib().flow2d_graph_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def level_residuals(self):
        """(rms_u, rms_v) per level of the last compute with params.report_residuals = 1, coarsest level first."""
        ru, rv, n = (C.c_double * MAX_LEVELS)(), (C.c_double * MAX_LEVELS)(), C.c_int()
        self._check(lib().flow2d_level_residuals(self._h, ru, rv, MAX_LEVELS, C.byref(n)))
        return [(ru[i], rv[i]) for i in range(n.value)]

    def level_times(self):
        """[(level_ms, solve_ms)] of the last compute with params.report_level_times = 1, coarsest level first."""
        a, b, n = (C.c_float * MAX_LEVELS)(), (C.c_float * MAX_LEVELS)(), C.c_int()
        self._check(lib().flow2d_level_times(self._h, a, b, MAX_LEVELS, C.byref(n)))
        return [(a[i], b[i]) for i in range(n.value)]

    def level_outer_iterations(self):
        """Outer iterations that ran per level of the last compute, coarsest level first (flow2d_level_outer_iterations)."""
        it, n = (C.c_int * MAX_LEVELS)(), C.c_int()
        self._check(lib().flow2d_level_outer_iterations(self._h, it, MAX_LEVELS, C.byref(n)))
        return [int(it[i]) for i in range(n.value)]

    def stage_residual(self, d_f0, d_f1w, d_u, d_v, d_du, d_dv, d_phi, d_ksi, w, h, hx, hy, params):
        ru, rv = C.c_double(), C.c_double()
        self._check(lib().flow2d_stage_residual(self._h, _ptr(d_f0), _ptr(d_f1w), _ptr(d_u), _ptr(d_v), _ptr(d_du), _ptr(d_dv),
                                                _ptr(d_phi), _ptr(d_ksi), w, h, hx, hy, C.byref(params), C.byref(ru), C.byref(rv)))
        return ru.value, rv.value

    def launch_counts(self):
        """Kernel launches of the last call, by kernel (flow2d_last_launch_counts)."""
        counts = (C.c_longlong * KERNEL_KINDS)()
        self._check(lib().flow2d_last_launch_counts(self._h, counts))
        return {lib().flow2d_kernel_kind_name(k).decode(): int(counts[k]) for k in range(KERNEL_KINDS) if counts[k]}

    # -- containers (torch is only the allocator) --
    def container(self, fill=None):
        import torch
        t = torch.empty((self.height, self.pitch), dtype=torch.float32, device="cuda:%d" % self.device)
        if fill is not None:
            t.fill_(fill)
        return t

    def to_container(self, arr, fill=float("nan")):
        """Dense (h, w) host array -> device container with the level in its top-left corner."""
        import torch
        a = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float32))
        t = self.container(fill)
        t[: a.shape[0], : a.shape[1]] = a.to(t.device)
        return t

    @staticmethod
    def from_container(t, w, h):
        return t[:h, :w].cpu().numpy().copy()

    # -- the hot path --
    def compute(self, f0, f1, params, out_u=None, out_v=None):
        """flow2d_compute: dense host arrays in, dense host arrays out (numpy or pinned torch CPU tensors)."""
        def host_ptr(x):
            if isinstance(x, np.ndarray):
                assert x.dtype == np.float32 and x.flags["C_CONTIGUOUS"] and x.shape == (self.height, self.width)
                return C.c_void_p(x.ctypes.data)
            assert not x.is_cuda and x.is_contiguous() and tuple(x.shape) == (self.height, self.width)
            return C.c_void_p(x.data_ptr())
        if out_u is None:
            out_u = np.empty((self.height, self.width), np.float32)
            out_v = np.empty((self.height, self.width), np.float32)
        self._check(lib().flow2d_compute(self._h, host_ptr(f0), host_ptr(f1), host_ptr(out_u), host_ptr(out_v), C.byref(params)))
        return out_u, out_v

    def compute_async(self, f0, f1, params, out_u, out_v):
        """flow2d_compute_async: pinned torch CPU tensors in/out; returns at once, finish with synchronize()."""
        for x in (f0, f1, out_u, out_v):
            assert not x.is_cuda and x.is_pinned() and x.is_contiguous() and tuple(x.shape) == (self.height, self.width)
        self._check(lib().flow2d_compute_async(self._h, C.c_void_p(f0.data_ptr()), C.c_void_p(f1.data_ptr()),
                                               C.c_void_p(out_u.data_ptr()), C.c_void_p(out_v.data_ptr()), C.byref(params)))

    def synchronize(self):
        self._check(lib().flow2d_synchronize(self._h))

    def prepare(self, params):
        """flow2d_prepare: capture the schedule of compute / compute_async for these parameters now."""
        self._check(lib().flow2d_prepare(self._h, C.byref(params)))

    def compute_device(self, d_f0, d_f1, d_u, d_v, params):
        self._check(lib().flow2d_compute_device(self._h, _ptr(d_f0), _ptr(d_f1), _ptr(d_u), _ptr(d_v), C.byref(params)))

    # -- one large frame on several GPUs (row slabs; see cuda_flow2d_b200.slab for the wiring helpers) --
    def slab_mailbox(self):
        """(device address, bytes) of this handle's mailbox (flow2d_slab_mailbox)."""
        ptr, n = C.c_void_p(), C.c_size_t()
        self._check(lib().flow2d_slab_mailbox(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def slab_export(self):
        buf = C.create_string_buffer(64)
        self._check(lib().flow2d_slab_export(self._h, buf))
        return buf.raw

    def slab_import(self, ipc_handle):
        ptr = C.c_void_p()
        self._check(lib().flow2d_slab_import(self._h, C.create_string_buffer(bytes(ipc_handle), 64), C.byref(ptr)))
        return ptr.value

    def slab_connect(self, rank, world, mailbox_above, mailbox_below, min_rows=0):
        self._check(lib().flow2d_slab_connect(self._h, rank, world, C.c_void_p(mailbox_above or 0), C.c_void_p(mailbox_below or 0), min_rows))

    def slab_rows(self, level_height=None):
        a, b = C.c_size_t(), C.c_size_t()
        self._check(lib().flow2d_slab_rows(self._h, level_height or self.height, C.byref(a), C.byref(b)))
        return a.value, b.value

    def slab_status(self):
        """Waits for the handle's stream and raises if a wait for a neighbour timed out (flow2d_slab_status)."""
        self._check(lib().flow2d_slab_status(self._h))

    def slab_stats(self):
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_int()
        self._check(lib().flow2d_slab_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"exchanges": a.value, "bytes_sent": b.value, "levels_slabbed": c.value}

    def compute_slab_device(self, d_f0, d_f1, d_u, d_v, params):
        """flow2d_compute_slab_device on a connected handle: this rank's rows (slab_rows()) of d_u, d_v are the result."""
        self._check(lib().flow2d_compute_slab_device(self._h, _ptr(d_f0), _ptr(d_f1), _ptr(d_u), _ptr(d_v), C.byref(params)))

    def stage_solve_slab(self, d_f0, d_f1w, d_u, d_v, d_du, d_dv, w, h, hx, hy, params):
        slabbed = C.c_int()
        self._check(lib().flow2d_stage_solve_slab(self._h, _ptr(d_f0), _ptr(d_f1w), _ptr(d_u), _ptr(d_v), _ptr(d_du), _ptr(d_dv),
                                                  w, h, hx, hy, C.byref(params), C.byref(slabbed)))
        return bool(slabbed.value)

    def debug_timing(self, d_stamps):
        self._check(lib().flow2d_debug_timing(self._h, _ptr(d_stamps)))

    # -- per-stage API on containers --
    def stage_blur(self, d_in, d_out, w, h, sigma):
        self._check(lib().flow2d_stage_blur(self._h, _ptr(d_in), _ptr(d_out), w, h, sigma))

    def stage_resample(self, d_in, iw, ih, d_out, ow, oh):
        self._check(lib().flow2d_stage_resample(self._h, _ptr(d_in), iw, ih, _ptr(d_out), ow, oh))

    def stage_warp(self, d_f0, d_f1, d_u, d_v, d_out, w, h, hx, hy):
        self._check(lib().flow2d_stage_warp(self._h, _ptr(d_f0), _ptr(d_f1), _ptr(d_u), _ptr(d_v), _ptr(d_out), w, h, hx, hy))

    def stage_solve(self, d_f0, d_f1w, d_u, d_v, d_du, d_dv, d_phi, d_ksi, w, h, hx, hy, params):
        self._check(lib().flow2d_stage_solve(self._h, _ptr(d_f0), _ptr(d_f1w), _ptr(d_u), _ptr(d_v), _ptr(d_du), _ptr(d_dv),
                                             _ptr(d_phi), _ptr(d_ksi), w, h, hx, hy, C.byref(params)))

    def stage_add(self, d_a, d_b, w, h):
        self._check(lib().flow2d_stage_add(self._h, _ptr(d_a), _ptr(d_b), w, h))

    def stage_median(self, d_in, d_out, w, h, radius):
        self._check(lib().flow2d_stage_median(self._h, _ptr(d_in), _ptr(d_out), w, h, radius))

    def stage_add_median(self, d_a, d_b, d_out, w, h, radius):
        self._check(lib().flow2d_stage_add_median(self._h, _ptr(d_a), _ptr(d_b), _ptr(d_out), w, h, radius))
