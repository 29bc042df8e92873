"""Row-slab decomposition of the solve (include/flow2d.h, BASELINE.json configs[4]) checked on ONE GPU:
N logical ranks = N handles in N threads with explicit ghost-row copies between barriers
(cuda_flow2d_b200.slab.ThreadedExchange).  Jacobi is tiling independent, so every rank must end with
exactly the bits of the single-handle result."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_():
    import torch
    assert torch.cuda.is_available()
    return torch


def _run_ranks(world, fn):
    out, err = [None] * world, []

    def work(r):
        try:
            out[r] = fn(r)
        except Exception as e:  # pragma: no cover
            err.append(e)
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    assert not err, err
    return out


@pytest.mark.parametrize("world,w,h,outer,inner,spp", [(2, 200, 520, 11, 5, 0), (3, 131, 800, 9, 5, 0), (2, 96, 600, 6, 12, 5), (4, 64, 1100, 10, 3, 0)])
def test_slabbed_solve_equals_single_handle(pkg, synth, torch_, world, w, h, outer, inner, spp):
    from cuda_flow2d_b200 import slab as S
    f0, f1, _, _ = synth.make_pair(w, h, 50 + w, U1=1.5, L=48.0)
    u = synth.smooth_random(w, h, 1, -2, 2)
    v = synth.smooth_random(w, h, 2, -2, 2)
    p = pkg.default_params(outer=outer, inner=inner, alpha=20.0, sweeps_per_pass=spp)
    ref = pkg.Flow2D(w, h)
    d = [ref.to_container(a, 0.0) for a in (f0, f1, u, v)]
    rdu, rdv = ref.container(0.0), ref.container(0.0)
    ref.stage_solve(d[0], d[1], d[2], d[3], rdu, rdv, None, None, w, h, 1.3, 1.1, p)
    torch_.cuda.synchronize()
    edu, edv = ref.from_container(rdu, w, h), ref.from_container(rdv, w, h)

    ex = S.ThreadedExchange(world, "cuda:0")

    def rank(r):
        fl = pkg.Flow2D(w, h)
        st = torch_.cuda.Stream()
        fl.set_stream(st.cuda_stream)
        dd = [fl.to_container(a, 0.0) for a in (f0, f1, u, v)]
        du, dv = fl.container(0.0), fl.container(0.0)
        torch_.cuda.synchronize()
        fl.stage_solve_slab(dd[0], dd[1], dd[2], dd[3], du, dv, w, h, 1.3, 1.1, p, ex.slab(r, min_rows=64))
        torch_.cuda.synchronize()
        return fl.from_container(du, w, h), fl.from_container(dv, w, h)

    res = _run_ranks(world, rank)
    assert ex.calls[1] == 1 and ex.calls[0] >= 1, ex.calls  # the level really was slabbed and exchanged
    for r, (gdu, gdv) in enumerate(res):
        assert np.array_equal(gdu, edu) and np.array_equal(gdv, edv), "rank %d differs" % r


def test_slabbed_flow_equals_single_handle(pkg, synth, torch_):
    from cuda_flow2d_b200 import slab as S
    w, h, world = 160, 700, 2
    f0, f1, _, _ = synth.make_pair(w, h, 9, U0=(0.5, -0.3), U1=0.8, L=64.0)
    p = pkg.default_params(levels=8, outer=10, inner=5, alpha=20.0)
    ref = pkg.Flow2D(w, h)
    eu, ev = ref.compute(f0, f1, p)
    ex = S.ThreadedExchange(world, "cuda:0")

    def rank(r):
        fl = pkg.Flow2D(w, h)
        st = torch_.cuda.Stream()
        fl.set_stream(st.cuda_stream)
        d0, d1 = fl.to_container(f0, 0.0), fl.to_container(f1, 0.0)
        du, dv = fl.container(0.0), fl.container(0.0)
        torch_.cuda.synchronize()
        fl.compute_slab_device(d0, d1, du, dv, p, ex.slab(r, min_rows=64))
        torch_.cuda.synchronize()
        return fl.from_container(du, w, h), fl.from_container(dv, w, h)

    res = _run_ranks(world, rank)
    assert ex.calls[1] >= 2, ex.calls  # several levels were slabbed
    for r, (gu, gv) in enumerate(res):
        assert np.array_equal(gu, eu) and np.array_equal(gv, ev), "rank %d differs" % r


def test_world_of_one_needs_no_transport(pkg, synth, torch_):
    from cuda_flow2d_b200 import slab as S
    w, h = 96, 80
    f0, f1, _, _ = synth.make_pair(w, h, 3)
    p = pkg.default_params(levels=4, outer=3)
    fl = pkg.Flow2D(w, h)
    eu, ev = fl.compute(f0, f1, p)
    d0, d1 = fl.to_container(f0, 0.0), fl.to_container(f1, 0.0)
    du, dv = fl.container(0.0), fl.container(0.0)
    fl.compute_slab_device(d0, d1, du, dv, p, S.Slab(0, 1, S.EXCHANGE_FN(0), None, 0))
    torch_.cuda.synchronize()
    assert np.array_equal(fl.from_container(du, w, h), eu) and np.array_equal(fl.from_container(dv, w, h), ev)
