"""Row-slab decomposition (include/flow2d.h, BASELINE.json configs[4]) checked on ONE GPU: N logical ranks = N handles
on the same device, one host thread and one stream each, wired mailbox to mailbox (cuda_flow2d_b200.slab.SlabGroup).
The same kernels, flags and row arithmetic as on N GPUs; only the stores stay on one device.  Jacobi is tiling
independent, so every rank must end with exactly the bits of the single-handle result on its own rows
(SURVEY.md section 4, item 4)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(240)]


@pytest.fixture(scope="module")
def torch_():
    import torch
    assert torch.cuda.is_available()
    return torch


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

    grp = S.SlabGroup(pkg, w, h, [0] * world, min_rows=64)

    def rank(r, fl):
        st = torch_.cuda.Stream()
        fl.set_stream(st.cuda_stream)
        dd = [fl.to_container(a, 0.0) for a in (f0, f1, u, v)]
        du, dv = fl.container(float("nan")), fl.container(float("nan"))
        torch_.cuda.current_stream().synchronize()  # NOT the device: another rank may already be waiting for this one
        slabbed = fl.stage_solve_slab(dd[0], dd[1], dd[2], dd[3], du, dv, w, h, 1.3, 1.1, p)
        fl.slab_status()
        return slabbed, fl.from_container(du, w, h), fl.from_container(dv, w, h), fl.slab_stats()

    res = grp.run(rank)
    for r, (slabbed, gdu, gdv, stats) in enumerate(res):
        assert slabbed and stats["exchanges"] >= 1, (r, stats)  # the level really was slabbed and exchanged
        y0, y1 = S.slab_rows(h, r, world)
        assert np.array_equal(gdu[y0:y1], edu[y0:y1]) and np.array_equal(gdv[y0:y1], edv[y0:y1]), "rank %d differs" % r
    grp.destroy()


@pytest.mark.parametrize("world,w,h,constancy,median", [(2, 160, 700, 0, 5), (3, 128, 900, 0, 7), (2, 160, 704, 1, 5)])
def test_slabbed_flow_equals_single_handle(pkg, synth, torch_, world, w, h, constancy, median):
    """Whole pyramid: prolongation, warp, derivatives (+ gradient tensor), robust iterations, add + median all on row
    slabs, the flow of a level handed to the next level through the neighbours' halo rows."""
    from cuda_flow2d_b200 import slab as S
    f0, f1, _, _ = synth.make_pair(w, h, 9, U0=(0.5, -0.3), U1=0.8, L=64.0)
    p = pkg.default_params(levels=8, outer=10, inner=5, alpha=20.0, median=median)
    ref = pkg.Flow2D(w, h, constancy=constancy)
    eu, ev = ref.compute(f0, f1, p)
    grp = S.SlabGroup(pkg, w, h, [0] * world, constancy=constancy, min_rows=64)

    def rank(r, fl):
        st = torch_.cuda.Stream()
        fl.set_stream(st.cuda_stream)
        d0, d1 = fl.to_container(f0, 0.0), fl.to_container(f1, 0.0)
        du, dv = fl.container(float("nan")), fl.container(float("nan"))
        torch_.cuda.current_stream().synchronize()  # NOT the device: another rank may already be waiting for this one
        fl.compute_slab_device(d0, d1, du, dv, p)
        fl.slab_status()
        return fl.from_container(du, w, h), fl.from_container(dv, w, h), fl.slab_stats()

    res = grp.run(rank)
    for r, (gu, gv, stats) in enumerate(res):
        assert stats["levels_slabbed"] >= 2, stats  # several levels were slabbed
        y0, y1 = S.slab_rows(h, r, world)
        assert np.array_equal(gu[y0:y1], eu[y0:y1]) and np.array_equal(gv[y0:y1], ev[y0:y1]), "rank %d differs" % r
    # a second flow through the same connected handles (epochs keep counting)
    res = grp.run(rank)
    for r, (gu, gv, stats) in enumerate(res):
        y0, y1 = S.slab_rows(h, r, world)
        assert np.array_equal(gu[y0:y1], eu[y0:y1]) and np.array_equal(gv[y0:y1], ev[y0:y1]), "rank %d differs (2nd call)" % r
    grp.destroy()


def test_world_of_one_needs_no_transport(pkg, synth, torch_):
    w, h = 96, 80
    f0, f1, _, _ = synth.make_pair(w, h, 3)
    p = pkg.default_params(levels=4, outer=3)
    fl = pkg.Flow2D(w, h)
    eu, ev = fl.compute(f0, f1, p)
    d0, d1 = fl.to_container(f0, 0.0), fl.to_container(f1, 0.0)
    du, dv = fl.container(0.0), fl.container(0.0)
    fl.compute_slab_device(d0, d1, du, dv, p)  # never connected: world = 1
    torch_.cuda.synchronize()
    assert np.array_equal(fl.from_container(du, w, h), eu) and np.array_equal(fl.from_container(dv, w, h), ev)
    assert fl.slab_rows() == (0, h)
