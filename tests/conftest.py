import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): oracle/liboracle.so through oracle/oracle.py."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def pkg():
    """The product package; the C-ABI library must already be built (no fallback)."""
    import flow2d_loader
    m = flow2d_loader.load()
    if not os.path.exists(m.lib_path()):
        m.build()
    return m


@pytest.fixture(scope="session")
def synth(pkg):
    from cuda_flow2d_b200 import synth as S
    return S


@pytest.fixture(scope="session")
def rub():
    """The reference's bundled frame pair (data/rub1.raw, rub2.raw: 8-bit 584x388),
    committed as a compressed fixture because /root/reference does not exist on the GPU box."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "rub_u8.npz"))
    return z["rub1"].astype(np.float32), z["rub2"].astype(np.float32)


REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def ref_available(variant=""):
    return os.path.exists(os.path.join(REF_DIR, variant, "ref_harness"))


class RefHarness:
    """Drives the reference's own CUDA build (oracle/_ref/ref_harness) one operator at a time."""

    def __init__(self, tmpdir, variant=""):
        self.exe = os.path.join(REF_DIR, variant, "ref_harness")
        self.tmp = str(tmpdir)
        self.n = 0

    def _w(self, arr, W, H):
        a = np.zeros((H, W), np.float32)
        a[: arr.shape[0], : arr.shape[1]] = arr
        self.n += 1
        p = os.path.join(self.tmp, "in%d.raw" % self.n)
        a.tofile(p)
        return p

    def _o(self):
        self.n += 1
        return os.path.join(self.tmp, "out%d.raw" % self.n)

    def _run(self, args):
        r = subprocess.run([self.exe] + [str(a) for a in args], stdin=subprocess.DEVNULL, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, timeout=600)
        if r.returncode != 0:
            raise RuntimeError("ref_harness %s failed (%d):\n%s" % (args[0], r.returncode, r.stdout.decode()[-2000:]))
        return r.stdout.decode()

    @staticmethod
    def _f(x):
        return "%.9g" % float(x)

    def _r(self, path, W, H, w, h):
        return np.fromfile(path, np.float32).reshape(H, W)[:h, :w].copy()

    def conv(self, img, sigma):
        H, W = img.shape
        o = self._o()
        self._run(["conv", W, H, self._f(sigma), self._w(img, W, H), o])
        return self._r(o, W, H, W, H)

    def resample(self, img, ow, oh, W=None, H=None):
        ih, iw = img.shape
        W, H = W or max(iw, ow), H or max(ih, oh)
        o = self._o()
        self._run(["resample", W, H, iw, ih, ow, oh, self._w(img, W, H), o])
        return self._r(o, W, H, ow, oh)

    def warp(self, f0, f1, u, v, hx, hy):
        h, w = f0.shape
        o = self._o()
        self._run(["warp", w, h, w, h, self._f(hx), self._f(hy)] + [self._w(a, w, h) for a in (f0, f1, u, v)] + [o])
        return self._r(o, w, h, w, h)

    def solve(self, f0, f1w, u, v, hx, hy, alpha, e_smooth, e_data, outer, inner, constancy=0, W=None, H=None):
        h, w = f0.shape
        W, H = W or w, H or h
        outs = [self._o() for _ in range(4)]
        self._run(["solve", W, H, w, h, self._f(hx), self._f(hy), self._f(alpha), self._f(e_smooth), self._f(e_data),
                   outer, inner, constancy] + [self._w(a, W, H) for a in (f0, f1w, u, v)] + outs)
        return tuple(self._r(o, W, H, w, h) for o in outs)  # du, dv, phi, ksi

    def add(self, a, b):
        h, w = a.shape
        o = self._o()
        self._run(["add", w, h, w, h, self._w(a, w, h), self._w(b, w, h), o])
        return self._r(o, w, h, w, h)

    def median(self, img, radius):
        h, w = img.shape
        o = self._o()
        self._run(["median", w, h, w, h, radius, self._w(img, w, h), o])
        return self._r(o, w, h, w, h)

    def flow(self, f0, f1, p, constancy=0, warmup=0, reps=1):
        """Whole reference ComputeFlow; p = dict(levels, scale, outer, inner, alpha, e_smooth, e_data, median, sigma)."""
        h, w = f0.shape
        a, b = self._w(f0, w, h), self._w(f1, w, h)
        pre = os.path.join(self.tmp, "flow%d_" % self.n)
        out = self._run(["flow", a, b, w, h, pre, p["levels"], self._f(p["scale"]), p["outer"], p["inner"], self._f(p["alpha"]),
                         self._f(p["e_smooth"]), self._f(p["e_data"]), p["median"], self._f(p["sigma"]), constancy, warmup, reps])
        ms = [float(l.split()[2]) for l in out.splitlines() if l.startswith("REF_MS timed")]
        u = np.fromfile(pre + "u.raw", np.float32).reshape(h, w)
        v = np.fromfile(pre + "v.raw", np.float32).reshape(h, w)
        return u, v, ms


@pytest.fixture
def ref(tmp_path):
    if not ref_available():
        pytest.skip("oracle/_ref (the reference's CUDA build) is not present; run `make -C oracle ref` where /root/reference exists")
    return RefHarness(tmp_path)


@pytest.fixture
def ref_cubin(tmp_path):
    if not ref_available("cubin"):
        pytest.skip("oracle/_ref/cubin is not present")
    return RefHarness(tmp_path, "cubin")


def epd_stats(u, v, ur, vr):
    """Endpoint difference statistics in pixels."""
    d = np.hypot(u.astype(np.float64) - ur, v.astype(np.float64) - vr)
    return {"mean": float(d.mean()), "p999": float(np.quantile(d, 0.999)), "max": float(d.max()),
            "n_gt_1e-2": int((d > 1e-2).sum()), "exact": bool(np.array_equal(u, ur) and np.array_equal(v, vr))}
