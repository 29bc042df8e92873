"""CPU checks of the thread-block-cluster solve (csrc/solve_cluster.cu): the block decomposition the scheduler picks, and the
index algebra of the kernel -- which shared-memory cell a thread owns, which four it reads, where an edge cell is pushed to in
the CTA next door -- exercised by running a stencil with the kernel's phase structure over those very indices
(flow2d_debug_cluster_cell calls csrc/solve_cluster_geom.h, the function the kernel calls) and comparing it with the same
stencil on the whole level with the reference's mirrored border (solve_2d.cu:243-262: -1 -> 1, n -> n-2).

The arithmetic of the kernel is one_px_outer, shared with solve_tiny / solve_small_pass; its bits are compared with the oracle
and the reference build on the GPU (tests/test_cluster_gpu.py)."""
import ctypes as C
import math

import numpy as np
import pytest


def _lib(pkg):
    L = C.CDLL(pkg.lib_path())
    L.flow2d_cluster_shape.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.flow2d_debug_cluster_cell.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    return L


def _shape(L, rw, rh, compact=0):
    out = (C.c_int * 5)()
    rc = L.flow2d_cluster_shape(rw, rh, compact, out)
    return rc, tuple(out)


def test_planner_invariants(pkg):
    L = _lib(pkg)
    rng = np.random.default_rng(5)
    sizes = [(2, 2), (32, 32), (33, 32), (128, 128), (129, 128), (16, 1024), (1024, 16), (8, 2048), (2048, 8), (2047, 9), (100, 90), (64, 64), (45, 23)]
    sizes += [tuple(int(v) for v in rng.integers(2, 200, 2)) for _ in range(300)]
    for rw, rh in sizes:
        for compact in (0, 1):
            rc, (cx, cy, tw, th, threads) = _shape(L, rw, rh, compact)
            if rc != 0:
                # no shape fits: even 16 CTAs of 1024 cells with blocks of at least 8 x 8 do not cover the region
                assert rc == -5
                assert all(max(8, math.ceil(rw / c)) * max(8, math.ceil(rh / (16 // c))) > 1024 for c in (1, 2, 4, 8, 16)), (rw, rh)
                continue
            assert cx * cy in (2, 4, 8, 16) and threads % 32 == 0 and threads <= 1024
            assert tw >= 8 and th >= 8 and tw * th <= threads < tw * th + 32
            assert cx * tw >= rw and cy * th >= rh
            # the shared planes of the kernel (sized for up to 256, 512 or 1024 cells) hold the block and its ring
            bucket = 256 if tw * th <= 256 else 512 if tw * th <= 512 else 1024
            assert (tw + 2) * (th + 2) <= bucket + bucket // 4 + 20
    # a region that fits is never refused; latency first (many small CTAs) by default, few CTAs when compact
    assert _shape(L, 64, 64)[1][4] == 256 and _shape(L, 64, 64)[1][0] * _shape(L, 64, 64)[1][1] == 16
    assert _shape(L, 64, 64, 1)[1][4] == 1024 and _shape(L, 64, 64, 1)[1][0] * _shape(L, 64, 64, 1)[1][1] == 4
    assert _shape(L, 128, 128)[1] == (4, 4, 32, 32, 1024)
    assert _shape(L, 128, 64)[1] == (4, 4, 32, 16, 512)
    assert _shape(L, 129, 128)[0] == -5


class _Cluster:
    """One cluster of the kernel on the CPU: per CTA a few shared planes with ring, threads = the kernel's index records."""

    NPLANES = 5  # U, PHI, W (unused as a plane), S0, S1

    def __init__(self, L, geom, level, cluster):
        self.cx, self.cy, self.tw, self.th, _ = geom
        self.ranks = self.cx * self.cy
        g = (C.c_int * 5)(*geom)
        lv = (C.c_int * 7)(*level)
        self.cells = []
        for r in range(self.ranks):
            row = []
            for t in range(self.tw * self.th):  # (threads without a cell only take part in the barriers)
                out = (C.c_int * 15)()
                assert L.flow2d_debug_cluster_cell(g, lv, r, cluster, t, out) == 0
                row.append(tuple(out))
            self.cells.append(row)
        self.plane = self.cells[0][0][14]
        assert (self.tw + 2) * (self.th + 2) <= self.plane
        self.mem = np.full((self.ranks, self.NPLANES, self.plane), np.nan)
        self.written = np.zeros((self.ranks, self.NPLANES, self.plane), bool)

    def publish(self, plane, values):
        """values[rank][t] -> own plane and, from edge cells, the ring of the CTA next door (pub<> of solve_onepx.cuh)"""
        targets = set()
        for r in range(self.ranks):
            for t, c in enumerate(self.cells[r]):
                ac, _, _, _, _, hr, ho, vr, vo = c[:9]
                v = values[r][t]
                for rank, off in ((r, ac), (hr, ho), (vr, vo)):
                    if rank < 0:
                        continue
                    assert 0 <= rank < self.ranks and 0 <= off < (self.tw + 2) * (self.th + 2)
                    assert (rank, off) not in targets, "two threads store into one cell"
                    targets.add((rank, off))
                    self.mem[rank, plane, off] = v
                    self.written[rank, plane, off] = True

    def neighbours(self, plane, r, t):
        c = self.cells[r][t]
        for off in c[1:5]:
            # every cell a thread reads has been written by its owner or pushed by the CTA next door
            assert self.written[r, plane, off], ("unwritten ring cell read", r, t, c)
        return [self.mem[r, plane, off] for off in c[1:5]]


def _mirror(i, n):
    return 1 if i < 0 else (n - 2 if i >= n else i)


def _global_reference(field, sweeps):
    """phi from the neighbours of the field, weights from the neighbours of phi, then Jacobi sweeps: the dependency
    structure of one outer iteration (solve_2d.cu:141-162, 333-367) with the mirrored border."""
    h, w = field.shape

    def nb(a, x, y):
        return a[y, _mirror(x - 1, w)], a[y, _mirror(x + 1, w)], a[_mirror(y - 1, h), x], a[_mirror(y + 1, h), x]

    phi = np.empty_like(field)
    for y in range(h):
        for x in range(w):
            l, r, u, d = nb(field, x, y)
            phi[y, x] = ((r - l) * 0.5 + (d - u) * 0.25) + field[y, x] * 0.125
    wgt = np.empty_like(field)
    for y in range(h):
        for x in range(w):
            l, r, u, d = nb(phi, x, y)
            wgt[y, x] = (((l + r) + u) + d) * 0.0625 + 0.5
    s = field.copy()
    for _ in range(sweeps):
        n = np.empty_like(s)
        for y in range(h):
            for x in range(w):
                l, r, u, d = nb(s, x, y)
                n[y, x] = (((l * 0.1 + r * 0.2) + u * 0.3) + d * 0.15) * wgt[y, x] + field[y, x] * 0.25
        s = n
    return s


def _run_cluster(L, geom, level, cluster, field, sweeps, out):
    h, w = field.shape
    cl = _Cluster(L, geom, level, cluster)
    U, PHI, S = 0, 1, (3, 4)
    own = [[field[min(max(c[10], 0), h - 1), min(max(c[9], 0), w - 1)] if c[12] else np.nan for c in row] for row in cl.cells]
    cl.publish(U, own)
    phi = [[None] * len(row) for row in cl.cells]
    for r, row in enumerate(cl.cells):
        for t in range(len(row)):
            l, rr, u, d = cl.neighbours(U, r, t)
            phi[r][t] = ((rr - l) * 0.5 + (d - u) * 0.25) + own[r][t] * 0.125
    cl.publish(PHI, phi)
    wgt = [[None] * len(row) for row in cl.cells]
    for r, row in enumerate(cl.cells):
        for t in range(len(row)):
            l, rr, u, d = cl.neighbours(PHI, r, t)
            wgt[r][t] = (((l + rr) + u) + d) * 0.0625 + 0.5
    cl.publish(S[0], own)
    cur = own
    for k in range(sweeps):
        new = [[None] * len(row) for row in cl.cells]
        for r, row in enumerate(cl.cells):
            for t in range(len(row)):
                l, rr, u, d = cl.neighbours(S[k % 2], r, t)
                new[r][t] = (((l * 0.1 + rr * 0.2) + u * 0.3) + d * 0.15) * wgt[r][t] + own[r][t] * 0.25
        cl.publish(S[(k + 1) % 2], new)
        cur = new
    n_out = 0
    for r, row in enumerate(cl.cells):
        for t, c in enumerate(row):
            if c[13]:
                assert c[12] and 0 <= c[9] < w and 0 <= c[10] < h
                assert np.isnan(out[c[10], c[9]]), "a pixel is produced twice"
                out[c[10], c[9]] = cur[r][t]
                n_out += 1
    return n_out


@pytest.mark.parametrize("w,h,compact", [(40, 30, 0), (33, 17, 0), (16, 9, 1), (100, 37, 0), (61, 64, 1), (2, 2, 0), (9, 120, 0)])
def test_whole_level_indices(pkg, w, h, compact):
    """the region covers the level: every pixel is produced once and equals the global stencil with the mirrored border"""
    L = _lib(pkg)
    rc, (cx, cy, tw, th, threads) = _shape(L, w, h, compact)
    assert rc == 0
    field = np.random.default_rng(w * 1000 + h).random((h, w))
    out = np.full((h, w), np.nan)
    n = _run_cluster(L, (cx, cy, tw, th, 1), (w, h, cx * tw, cy * th, 0, 0, h), 0, field, 4, out)
    assert n == w * h
    assert np.array_equal(out, _global_reference(field, 4))


@pytest.mark.parametrize("w,h,geom,sweeps", [(70, 50, (2, 2, 16, 16), 2), (45, 90, (4, 2, 8, 16), 3), (31, 33, (2, 4, 16, 8), 1)])
def test_pass_mode_indices(pkg, w, h, geom, sweeps):
    """a grid of clusters with an S+1 halo: the output tiles cover the level once and are exact after S sweeps"""
    L = _lib(pkg)
    cx, cy, tw, th = geom
    halo = sweeps + 1
    ow, oh = cx * tw - 2 * halo, cy * th - 2 * halo
    assert ow > 0 and oh > 0
    ncx, ncy = -(-w // ow), -(-h // oh)
    field = np.random.default_rng(w + h).random((h, w))
    out = np.full((h, w), np.nan)
    total = 0
    for cl in range(ncx * ncy):
        total += _run_cluster(L, (cx, cy, tw, th, ncx), (w, h, ow, oh, halo, 0, h), cl, field, sweeps, out)
    assert total == w * h
    assert np.array_equal(out, _global_reference(field, sweeps))


def test_pass_mode_row_range(pkg):
    """y0, y1: only those rows of the level are produced (the pass of a level whose other rows are not needed)"""
    L = _lib(pkg)
    w, h, sweeps = 40, 60, 2
    cx, cy, tw, th = 2, 2, 16, 16
    halo = sweeps + 1
    ow, oh = cx * tw - 2 * halo, cy * th - 2 * halo
    y0, y1 = 13, 47
    ncx, ncy = -(-w // ow), -(-(y1 - y0) // oh)
    field = np.random.default_rng(3).random((h, w))
    out = np.full((h, w), np.nan)
    for cl in range(ncx * ncy):
        _run_cluster(L, (cx, cy, tw, th, ncx), (w, h, ow, oh, halo, y0, y1), cl, field, sweeps, out)
    ref = _global_reference(field, sweeps)
    assert np.array_equal(out[y0:y1], ref[y0:y1])
    assert np.isnan(out[:y0]).all() and np.isnan(out[y1:]).all()
