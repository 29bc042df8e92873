#!/usr/bin/env python
"""bench.py -- throughput of the optical-flow hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c4|c2|rub_c1b|rub_c1a|c1b|c3|c3g|c5] [--streams K] [--pairs P] [--no-extra]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (flow2d_compute*) over one batch of frame pairs per GPU.  Default
workload = BASELINE.json configs[3] ("c4"): batch of synthetic 1024x1024 radiography-like pairs through the
FULL model (50-level pyramid, presmoothing, warping, 40x5 robust iterations, median) -- the largest config
of the metric that fits one GPU and the one its "1/2/4/8 B200" is quoted on (pairs sharded across the GPUs,
weak scaling, no data-path collective).  The other configs ride along in `extra` (C2: single level, 500
sweeps; C3: 2048x2048 gradient constancy; C5: 8192x8192 on one GPU, or slabbed by rows across the N GPUs =
strong scaling, with its exchange bytes), each with its own device-timed and end-to-end number.
Rank 0 prints ONE JSON line.

  value      whole-job Mpix/s with the frames already resident in HBM (flow2d_compute_device),
             timed with CUDA events on the launching streams, max over ranks
  e2e        same metric through flow2d_compute_async + flow2d_synchronize with PINNED HOST buffers:
             H2D of both frames and D2H of both flow fields inside the timed region of every step
  roofline   the kernel with the largest share of the step (by live per-level timings): algorithmic
             bytes per launch / measured launch duration against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (a port of the reference algorithm) on the host cores, bounded sample
  --impl reference   the reference's own CUDA build (oracle/_ref/ref_harness; it has no CPU path), one
             process per GPU; falls back to the CPU oracle port when oracle/_ref is not present
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpix/s of converged flow (fixed settings)"
FULL = dict(levels=50, scale=0.9, outer=40, inner=5, alpha=35.0, e_smooth=0.001, e_data=0.001, median=5, sigma=1.5)

# SURVEY.md section 8(d)
WORKLOADS = {
    "c2": dict(name="C2: synthetic 1024x1024 f32 pair, Grey, single level, Horn-Schunck-style 500 Jacobi sweeps "
                    "(BASELINE.json configs[1])",
               w=1024, h=1024, seed=1001, gen=dict(U0=(0.3, -0.2), U1=0.5, L=256.0),
               cfg=dict(levels=1, scale=0.5, outer=1, inner=500, alpha=0.25, e_smooth=1.0, e_data=1000.0, median=1, sigma=0.0)),
    "c1b": dict(name="C1b-shaped: synthetic 584x388 pair, reference main.cpp defaults (47 levels, 40x5, median 5, sigma 1.5)",
                w=584, h=388, seed=1101, gen=dict(U0=(0.5, -0.3), U1=1.0, L=128.0), cfg=dict(FULL)),
    "rub_c1b": dict(name="C1b: the reference's bundled rub1/rub2 pair (584x388, 8-bit -> f32) with its main.cpp defaults "
                         "(47 levels, 40x5, alpha 35, median 5, sigma 1.5) (BASELINE.json configs[0])", rub=True,
                    w=584, h=388, seed=0, gen={}, cfg=dict(FULL)),
    "rub_c1a": dict(name="C1a: the bundled rub pair with the solver values of the reference's settings.xml "
                         "(20 levels, 20x5, alpha 3.5, median 5, sigma 0.45) (BASELINE.json configs[0])", rub=True,
                    w=584, h=388, seed=0, gen={},
                    cfg=dict(levels=20, scale=0.9, outer=20, inner=5, alpha=3.5, e_smooth=0.001, e_data=0.001, median=5, sigma=0.45)),
    "c4": dict(name="C4: batch of 1024x1024 radiography-like pairs (contrast 0.3, noise sigma 2), full model with the reference's "
                    "main.cpp defaults (50 levels, scale 0.9, 40x5 iterations, alpha 35, median 5, sigma 1.5), pairs sharded "
                    "across the GPUs (BASELINE.json configs[3])", streams=8, pairs=16,
               w=1024, h=1024, seed=4000, gen=dict(U0=(0.0, 0.0), U1=4.0, L=384.0, contrast=0.3, noise=2.0), cfg=dict(FULL)),
    "c3": dict(name="C3-Grey: synthetic 2048x2048 pair, full pyramid (50 levels, 40x5, median 5, sigma 1.5), alpha 3.5",
               w=2048, h=2048, seed=2001, gen=dict(U0=(3.0, -2.0), U1=6.0, L=512.0), cfg=dict(FULL, alpha=3.5)),
    "c3g": dict(name="C3: synthetic 2048x2048 pair, full model: GRADIENT constancy + robust penalisers + flow-driven smoothness, "
                     "full pyramid (50 levels, 40x5, median 5, sigma 1.5), alpha 3.5 (BASELINE.json configs[2])", gradient=True,
                w=2048, h=2048, seed=2001, gen=dict(U0=(3.0, -2.0), U1=6.0, L=512.0), cfg=dict(FULL, alpha=3.5)),
    "c5": dict(name="C5-Grey: single 8192x8192 pair, full pyramid (50 levels, 40x5, median 5, sigma 1.5), alpha 3.5; "
                    "rows slabbed across the GPUs (BASELINE.json configs[4])", slab=True,
               w=8192, h=8192, seed=5001, gen=dict(U0=(0.0, 0.0), U1=8.0, L=2048.0), cfg=dict(FULL, alpha=3.5)),
}


def config_of(wl):
    """The `config` object of the JSON line: identical, key for key and value for value, in both arms."""
    return {"workload": wl["name"], "frame": "%dx%d f32" % (wl["w"], wl["h"]), "settings": dict(wl["cfg"]),
            "data_term": "gradient" if wl.get("gradient") else "grey",
            "l2": "L2 flushed (256 MB write) between timed iterations"}


SMI_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + SMI_QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the committed
    `ncu --set full` capture (profiles/r02/<tag>_ncu_summary.txt, written by tools/gpu_round.sh), in bytes;
    None if there is no capture."""
    total, unit = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for rnd in ("r02", "r01"):
        path = os.path.join(ROOT, "profiles", rnd, "%s_ncu_summary.txt" % tag)
        if not os.path.exists(path):
            continue
        try:
            for line in open(path):
                f = line.split()
                if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(f[2].strip("[]'\"")) * unit[f[1]]
        except Exception:
            return None
        return total or None
    return None


def make_frames(wl, rank):
    if wl.get("rub"):  # the reference's own frame pair, committed as a fixture (tests/golden/README.md)
        z = np.load(os.path.join(ROOT, "tests", "golden", "rub_u8.npz"))
        return z["rub1"].astype(np.float32), z["rub2"].astype(np.float32)
    from cuda_flow2d_b200 import synth
    g = dict(wl["gen"])
    return synth.make_pair(wl["w"], wl["h"], wl["seed"] + rank, **g)[:2]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------
def run_reference(args, wl, rank, world):
    """The reference's own implementation of the path.  It is CUDA-only (no CPU path) and hard-wired
    to device 0 (cuda_operation_solve_2d.cpp:168), so at N GPUs rank 0 starts N processes of the unmodified reference
    build, one per GPU (CUDA_VISIBLE_DEVICES), each looping OpticalFlow2D::ComputeFlow over its own pair."""
    if rank != 0:
        return
    n = max(1, args.gpus)
    cfg, w, h = wl["cfg"], wl["w"], wl["h"]
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    line = {"impl": "reference", "metric": METRIC, "unit": "Mpix/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(wl),
            "schedule": {"pairs_per_step_per_gpu": 1, "concurrent_streams_per_gpu": 1,
                         "sharding": "one reference process per GPU, one frame pair per step each"}}
    if os.path.exists(exe):
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        ids = [x for x in visible.split(",") if x != ""] if visible else [str(i) for i in range(n)]
        ids = (ids * n)[:n]
        rcs = []
        with tempfile.TemporaryDirectory() as tmp:
            procs = []
            for g in range(n):
                f0, f1 = make_frames(wl, g * 16)
                a, b = os.path.join(tmp, "f0_%d.raw" % g), os.path.join(tmp, "f1_%d.raw" % g)
                f0.tofile(a)
                f1.tofile(b)
                cmd = [exe, "flow", a, b, w, h, "-", cfg["levels"], "%.9g" % cfg["scale"], cfg["outer"], cfg["inner"],
                       "%.9g" % cfg["alpha"], "%.9g" % cfg["e_smooth"], "%.9g" % cfg["e_data"], cfg["median"], "%.9g" % cfg["sigma"],
                       1 if wl.get("gradient") else 0, args.warmup, args.steps]
                procs.append([str(c) for c in cmd])

            def launch(cmds):
                ps = [subprocess.Popen(c, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=ids[g])) for g, c in enumerate(cmds)]
                outs = [p.communicate()[0].decode() for p in ps]
                return ([p.returncode for p in ps],
                        [[float(l.split()[2]) for l in o.splitlines() if l.startswith("REF_MS timed")] for o in outs])

            rcs, ms = launch(procs)
            if any(rcs) or not all(ms):
                # the reference build sometimes dies on a repeated ComputeFlow call (seen: SIGSEGV on the C4 frames);
                # fall back to one fresh process per step (Initialize / PTX JIT stay outside its timer)
                ms = [[] for _ in range(n)]
                once = [c[:-2] + ["0", "1"] for c in procs]
                for _ in range(max(1, args.steps)):
                    rcs, m1 = launch(once)
                    for g in range(n):
                        ms[g] += m1[g]
        if not all(ms):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness failed with codes %r" % (rcs,)}))
            return
        t = max(sum(x) / len(x) for x in ms)  # the slowest GPU's mean step
        v = n * w * h / (t * 1e-3) / 1e6
        line.update(value=v, ms_per_step=t, gpu_launches=None,
                    cpu_baseline={"value": v, "unit": "Mpix/s", "cores": n, "kind": "reference",
                                  "sample": "the reference's own CUDA build (it has no CPU path), one process and one host thread per "
                                            "GPU, full workload, wall clock around OpticalFlow2D::ComputeFlow incl. its H2D/D2H"},
                    e2e={"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 2 * w * h * 4 * n, "d2h_bytes_per_step": 2 * w * h * 4 * n})
    else:
        # no reference build on this box: time the CPU port of the same algorithm on all host cores
        from oracle import oracle as O
        O.set_num_threads(host_cores())
        f0, f1 = make_frames(wl, 0)
        p = O.make_params(constancy=1 if wl.get("gradient") else 0, **cfg)
        for _ in range(min(args.warmup, 1)):
            O.compute_flow(f0, f1, p)
        k = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(k):
            O.compute_flow(f0, f1, p)
        t = (time.perf_counter() - t0) / k * 1e3
        v = w * h / (t * 1e-3) / 1e6
        line.update(value=v, ms_per_step=t, steps=k, gpu_launches=0,
                    cpu_baseline={"value": v, "unit": "Mpix/s", "cores": O.num_threads(), "kind": "port",
                                  "sample": "CPU oracle (port), full workload, %d repetition(s)" % k},
                    e2e={"value": v, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def measure(ctx, key, steps, warmup, streams=0, pairs=0, detail=True, sample_clocks=False):
    """Times one workload on this rank's GPU (device-resident and end-to-end) and reduces over the ranks."""
    torch, m, dist, rank, world, dev = ctx.torch, ctx.m, ctx.dist, ctx.rank, ctx.world, ctx.dev
    wl = WORKLOADS[key]
    w, h, cfg = wl["w"], wl["h"], wl["cfg"]
    K = max(1, streams or wl.get("streams", 1))      # concurrent handles (one stream each) per GPU
    P = max(K, pairs or wl.get("pairs", K))          # frame pairs per step per GPU
    n_distinct = min(P, 4)
    slab_mode = bool(wl.get("slab"))
    if slab_mode:
        from cuda_flow2d_b200 import synth, slab as slab_mod
        g0, g1 = synth.make_pair_torch(w, h, wl["seed"], "cuda:%d" % dev, **wl["gen"])  # identical on every rank
        frames = [(g0.cpu().numpy(), g1.cpu().numpy())]
        del g0, g1
        K = P = n_distinct = 1
    else:
        frames = [make_frames(wl, rank * 16 + i) for i in range(n_distinct)]
    f0, f1 = frames[0]
    handles = [m.Flow2D(w, h, constancy=m.GRADIENT if wl.get("gradient") else m.GREY, device=dev) for _ in range(K)]
    fl = handles[0]
    params = m.default_params(**cfg)
    if os.environ.get("FLOW2D_BENCH_THROUGHPUT_MODE") is not None:  # A/B switch for measurements
        params.throughput_mode = int(os.environ["FLOW2D_BENCH_THROUGHPUT_MODE"])
    base = torch.cuda.Stream(device=dev)
    streams_ = [torch.cuda.Stream(device=dev) for _ in range(K)]
    for hd, st in zip(handles, streams_):
        hd.set_stream(st.cuda_stream)
    stream = streams_[0]

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def fan_out_in(enqueue):
        """start event on `base`, every worker stream waits for it, work is enqueued, `base` waits for all."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(base)
        for st in streams_:
            st.wait_event(a)
        enqueue()
        for st in streams_:
            e = torch.cuda.Event()
            e.record(st)
            base.wait_event(e)
        b.record(base)
        return a, b

    # ---- device-resident arm: frames already in HBM ----
    din = [(fl.to_container(frames[i % n_distinct][0], 0.0), fl.to_container(frames[i % n_distinct][1], 0.0)) for i in range(P)]
    dout = [(hd.container(0.0), hd.container(0.0)) for hd in handles]
    if ctx.flush is None:
        ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % dev)  # > 126 MB L2
    flush = ctx.flush

    transport = None
    if slab_mode and world > 1:
        # mailboxes wired through CUDA IPC handles; the halo rows then travel as peer stores + flags (csrc/slab.cu)
        slab_mod.connect_ipc(fl, dist, rank, world, "cuda:%d" % dev)
        transport = True
    own0, own1 = fl.slab_rows() if slab_mode else (0, h)

    def step_device():
        if transport is not None:
            fl.compute_slab_device(din[0][0], din[0][1], dout[0][0], dout[0][1], params)
            return
        for i in range(P):
            k = i % K
            handles[k].compute_device(din[i][0], din[i][1], dout[k][0], dout[k][1], params)

    for _ in range(warmup):
        step_device()
    barrier()
    launches_per_step = sum(hd.stats()["kernel_launches"] for hd in handles) * (P // K)
    launches_by_kernel = {}
    for hd in handles:
        for k, v in hd.launch_counts().items():
            launches_by_kernel[k] = launches_by_kernel.get(k, 0) + v * (P // K)
    sampler = ClockSampler(dev) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    ev = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        with torch.cuda.stream(base):
            flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
        ev.append(fan_out_in(step_device))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- end-to-end arm: pinned host buffers through the public host API ----
    hin = [(torch.from_numpy(frames[i % n_distinct][0]).pin_memory(), torch.from_numpy(frames[i % n_distinct][1]).pin_memory())
           for i in range(P)]
    hout = [(torch.empty((h, w), dtype=torch.float32).pin_memory(), torch.empty((h, w), dtype=torch.float32).pin_memory())
            for _ in range(P)]

    def step_e2e():
        if transport is not None:
            # public API of the slab path works on device containers: the copies are the caller's.  Every rank uploads
            # both frames (the warp reads rows it cannot know in advance) and downloads its own rows of the flow.
            with torch.cuda.stream(streams_[0]):
                din[0][0][:h, :w].copy_(hin[0][0], non_blocking=True)
                din[0][1][:h, :w].copy_(hin[0][1], non_blocking=True)
                fl.compute_slab_device(din[0][0], din[0][1], dout[0][0], dout[0][1], params)
                hout[0][0][own0:own1].copy_(dout[0][0][own0:own1, :w], non_blocking=True)
                hout[0][1][own0:own1].copy_(dout[0][1][own0:own1, :w], non_blocking=True)
            return
        for i in range(P):
            handles[i % K].compute_async(hin[i][0], hin[i][1], params, hout[i][0], hout[i][1])

    for _ in range(warmup):
        step_e2e()
        for hd in handles:
            hd.synchronize()
    barrier()
    ev2, t0 = [], time.perf_counter()
    for _ in range(steps):
        ev2.append(fan_out_in(step_e2e))
        for hd in handles:
            hd.synchronize()   # the step's result is on the host
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = sum(a.elapsed_time(b) for a, b in ev2)  # CUDA events around H2D .. D2H of the whole step
    clocks = sampler.stop() if sampler else None
    hu, hv = hout[0]
    result_check = float(hu.abs().mean() + hv.abs().mean())
    d0, d1 = din[0]

    roof = warp_roof = shares = None
    if detail and rank == 0:
        roof, warp_roof, shares = kernel_detail(ctx, wl, key, fl, stream, d0, d1, dev_ms / steps / P)

    # ---- reduce over ranks ----
    if dist is not None:
        tt = torch.tensor([dev_ms, e2e_ms, t_wall, e2e_wall], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, t_wall, e2e_wall = tt.tolist()

    pix = w * h * steps * P * (1 if slab_mode else world)
    slab_stats, n_slab_calls = (fl.slab_stats(), max(1, 2 * (steps + warmup))) if transport is not None else (None, 1)
    res = {
        "value": pix / (dev_ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": dev_ms / steps, "scaling": "strong" if slab_mode else "weak",
        "config": config_of(wl),
        "schedule": {"pairs_per_step_per_gpu": P, "concurrent_streams_per_gpu": K,
                     "sharding": ("rows of the %d largest levels slabbed across the GPUs, every stage on the rank's own rows; only "
                                  "halo rows move (peer stores + flags over NVLink, no collective): %d exchanges and %.1f MB sent "
                                  "per flow by rank 0" % (slab_stats["levels_slabbed"], slab_stats["exchanges"] // n_slab_calls,
                                                          slab_stats["bytes_sent"] / 1e6 / n_slab_calls))
                     if transport is not None else "pair i -> GPU i mod N, one handle per pair in flight, no data-path collective"},
        "e2e": {"value": pix / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": 2 * w * h * 4 * P,
                "d2h_bytes_per_step": 2 * w * (own1 - own0) * 4 * P, "ms_per_step": e2e_ms / steps,
                "wall_ms_per_step": e2e_wall / steps * 1e3, "api": "flow2d_compute_async + flow2d_synchronize (pinned host in/out)",
                "result_check": result_check},
        "gpu_launches": int(launches_per_step * steps),
        "wall_ms_per_step": t_wall / steps * 1e3,
        "launches_by_kernel": {k: int(v * steps) for k, v in launches_by_kernel.items()},
    }
    if clocks is not None:
        res["clocks"] = {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")}
    if roof is not None:
        res.update(roofline=roof, roofline_warp=warp_roof, kernel_time_shares=shares)
    for hd in handles:
        hd.destroy()
    del din, dout, hin, hout
    torch.cuda.empty_cache()
    return res, (f0, f1)


def kernel_detail(ctx, wl, key, fl, stream, d0, d1, ms_per_pair):
    """Live per-level timings of the solve (CUDA events on the launching stream) -> which kernel dominates the step,
    its launch duration at the finest level it runs on, and the roofline objects."""
    torch, m, dev = ctx.torch, ctx.m, ctx.dev
    w, h, cfg = wl["w"], wl["h"], wl["cfg"]
    sp = m.default_params(**cfg)
    t = [fl.container(0.0) for _ in range(4)]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_kind_ms, finest = {}, {}
    table = m.level_table(w, h, cfg["scale"], cfg["levels"])  # coarsest first
    for cw, ch, hx, hy in table:
        reps = 2 if cw * ch >= 1 << 16 else 4
        c0 = fl.launch_counts()
        with torch.cuda.stream(stream):
            fl.stage_solve(d0, d1, t[0], t[1], t[2], t[3], None, None, cw, ch, float(hx), float(hy), sp)  # warm-up
            c1 = fl.launch_counts()
            a.record(stream)
            for _ in range(reps):
                fl.stage_solve(d0, d1, t[0], t[1], t[2], t[3], None, None, cw, ch, float(hx), float(hy), sp)
            b.record(stream)
        torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b) / reps
        used = {k: v - c0.get(k, 0) for k, v in c1.items() if k.startswith("solve") and v > c0.get(k, 0)}
        if not used:
            continue
        kname, n_pass = max(used.items(), key=lambda kv: kv[1])  # launches of one solve of this level
        per_kind_ms[kname] = per_kind_ms.get(kname, 0.0) + ms
        finest[kname] = (cw, ch, ms, n_pass)  # the table ends with the finest level
    solve_ms = sum(per_kind_ms.values())
    shares = {k: round(v / solve_ms, 4) for k, v in per_kind_ms.items()}
    shares["ms_one_pair_alone"] = {k: round(v, 3) for k, v in per_kind_ms.items()}
    shares["note"] = ("share of the solve time of ONE pair run alone (%.3f ms; the solve is 88-99 %% of a flow), timed level by level "
                      "through flow2d_stage_solve with CUDA events; the step itself overlaps %s pairs on their own streams and takes "
                      "%.3f ms per pair" % (solve_ms, wl.get("streams", 1), ms_per_pair))
    kname = max(per_kind_ms.items(), key=lambda kv: kv[1])[0]
    cw, ch, ms, n_pass = finest[kname]
    launch_ms = ms / n_pass
    peak, peak_src = measured_peak_gbs()
    sweeps = cfg["outer"] * cfg["inner"] / n_pass
    # a fused phi/ksi + T-sweep pass reads 6 and writes 2 full-size fields = 32 B per level pixel
    # (a reference-shaped single sweep: 8R + 2W = 40 B).  C2's passes are later passes of ONE outer iteration and
    # move 9R + 2W; they are rated at the reference-shaped 40 B like in round 1.
    bpp = 40.0 if cfg["outer"] == 1 else 32.0
    alg_bytes = bpp * cw * ch
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    unfused = (32.0 + 40.0 * sweeps) if bpp == 32.0 else 40.0 * sweeps
    # the tiled pass has three generations (csrc/solve.cu, solve_pass2.cu, solve_pass3.cu); name the one that ran
    gen = kname.replace("(resident)", "")
    if kname == "solve_pass":
        gen = ("solve_pass" if os.environ.get("FLOW2D_SOLVE_V1") else
               "solve_pass2" if (wl.get("gradient") or os.environ.get("FLOW2D_SOLVE_V2")) else "solve_pass3")
    roof = {"kernel": "%s_kernel<%s> at the finest level it runs on (%dx%d; %d launches per solve, %.3g Jacobi sweeps per launch)" %
                      (gen, "gradient" if wl.get("gradient") else "grey", cw, ch, n_pass, sweeps),
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic("solve_" + key), "peak_source": peak_src, "launch_us": launch_ms * 1e3,
            "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_per_pixel": bpp,
            "share_of_step": shares.get(kname),
            "unfused_equivalent_gbs": unfused * cw * ch / (launch_ms * 1e-3) / 1e9,
            "note": "temporally blocked: one launch = robust weights + %g Jacobi sweeps, which the reference's decomposition moves "
                    "%g B/px for; unfused_equivalent_gbs = bytes of that decomposition over the same time" % (sweeps, unfused)}
    # second kernel BASELINE.json's metric names: backward warping, 16 B per pixel (u, v, frame 1 read; warped frame
    # written); several buffer sets, together larger than twice the L2, launched back to back: every byte comes from
    # HBM and the launch overhead is amortised as it is inside a pyramid
    g = m.level_geometry(w, h, cfg["scale"], 0)
    nsets = max(2, min(24, int(math.ceil(300e6 / (16.0 * fl.pitch * h)))))
    sets = [[fl.container(0.0) for _ in range(4)] for _ in range(nsets)]
    yy, xx = torch.meshgrid(torch.arange(h, device=d1.device, dtype=torch.float32),
                            torch.arange(fl.pitch, device=d1.device, dtype=torch.float32), indexing="ij")
    for q in sets:  # a smooth +-2 px flow, the magnitude of one pyramid level's update
        q[0].copy_(d1)
        q[1].copy_(2.0 * torch.sin(xx / 40.0) * torch.cos(yy / 50.0))
        q[2].copy_(2.0 * torch.cos(xx / 30.0) * torch.sin(yy / 60.0))
    del xx, yy
    with torch.cuda.stream(stream):
        for q in sets:  # warm-up
            fl.stage_warp(d0, q[0], q[1], q[2], q[3], w, h, float(g[2]), float(g[3]))
        a.record(stream)
        for _ in range(3):
            for q in sets:
                fl.stage_warp(d0, q[0], q[1], q[2], q[3], w, h, float(g[2]), float(g[3]))
        b.record(stream)
    torch.cuda.synchronize(dev)
    warp_ms = a.elapsed_time(b) / (3 * nsets)
    del sets
    warp_roof = {"kernel": "warp_kernel", "bound": "hbm", "achieved": 16.0 * w * h / (warp_ms * 1e-3) / 1e9, "peak": peak,
                 "unit": "GB/s", "frac": 16.0 * w * h / (warp_ms * 1e-3) / 1e9 / peak, "launch_us": warp_ms * 1e3,
                 "algorithmic_bytes_per_launch": 16.0 * w * h,
                 "note": "finest level of the workload; %d buffer sets (%.0f MB, > 2 x L2) warped back to back so that "
                         "every byte comes from HBM" % (nsets, nsets * 16.0 * fl.pitch * h / 1e6)}
    return roof, warp_roof, shares


def cpu_baseline(wl, f0, f1):
    """The CPU oracle on ALL host cores (thread count pinned explicitly: torchrun exports OMP_NUM_THREADS=1)."""
    from oracle import oracle as O
    O.set_num_threads(host_cores())
    w, h, cfg = wl["w"], wl["h"], wl["cfg"]
    c = dict(cfg)
    sample = "one pair, full workload once"
    scale = 1.0
    if c["outer"] * c["inner"] * w * h * (5.3 if c["levels"] > 1 else 1.0) > 3e9:
        # bound the sample: fewer outer iterations of the same workload, scaled linearly
        scale = c["outer"] / 4.0
        c["outer"] = 4
        sample = "one pair, outer iterations cut to 4 of %d (same pyramid), time scaled x%.1f" % (cfg["outer"], scale)
    p = O.make_params(constancy=1 if wl.get("gradient") else 0, **c)
    t0 = time.perf_counter()
    O.compute_flow(f0, f1, p)
    tc = (time.perf_counter() - t0) * scale
    return {"value": w * h / tc / 1e6, "unit": "Mpix/s", "cores": O.num_threads(), "kind": "port",
            "sample": "CPU oracle (plain-C port of the reference algorithm, OpenMP pinned to %d threads = all host cores); %s" %
                      (O.num_threads(), sample)}


def run_ours(args, wl_key, rank, world, local_rank):
    import torch
    import flow2d_loader
    m = flow2d_loader.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    ctx = Ctx()
    ctx.torch, ctx.m, ctx.rank, ctx.world, ctx.dev, ctx.dist, ctx.flush = torch, m, rank, world, local_rank, None, None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ctx.dist = dist
    torch.cuda.set_device(local_rank)
    wl = WORKLOADS[wl_key]
    res, (f0, f1) = measure(ctx, wl_key, args.steps, args.warmup, args.streams, args.pairs, detail=True, sample_clocks=True)

    extra = {}
    if not args.no_extra and wl_key == "c4":
        # every other config of BASELINE.json gets a driver-side record too (short runs)
        plan = [("c2", 10, 3), ("c3g", 4, 3), ("c5", 2, 1)] if world == 1 else [("c5", 2, 1)]
        for k, st, wu in plan:
            try:
                r, _ = measure(ctx, k, st, wu, detail=(world == 1 and k == "c2"))
                extra[k] = {x: r[x] for x in ("value", "unit", "ms_per_step", "scaling", "steps", "warmup", "config", "schedule",
                                              "e2e", "gpu_launches", "roofline", "roofline_warp") if x in r}
            except Exception as e:  # an extra must never take the headline down
                extra[k] = {"error": repr(e)}

    cpu = None
    if rank == 0 and wl.get("slab"):
        cpu = {"value": None, "unit": "Mpix/s", "cores": 0, "kind": "port",
               "sample": "not run for the 8192x8192 frame (minutes of CPU time); see the c3 workload for the same settings at 2048x2048"}
    elif rank == 0:
        cpu = cpu_baseline(wl, f0, f1)
    if rank == 0:
        line = {"metric": METRIC, "value": res["value"], "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": res["scaling"], "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": res["config"], "schedule": res["schedule"], "e2e": res["e2e"],
                "gpu_launches": res["gpu_launches"], "wall_ms_per_step": res["wall_ms_per_step"], "clocks": res.get("clocks"),
                "launches_by_kernel": res["launches_by_kernel"], "roofline": res.get("roofline"),
                "roofline_warp": res.get("roofline_warp"), "kernel_time_shares": res.get("kernel_time_shares"),
                "cpu_baseline": cpu, "extra": extra}
        print(json.dumps(line))
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="concurrent handles (streams) per GPU; 0 = workload default")
    ap.add_argument("--pairs", type=int, default=0, help="frame pairs per step per GPU; 0 = workload default")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra workloads (c2, c3g, c5) of the default run")
    args = ap.parse_args()
    if args.impl == "ours" and not WORKLOADS[args.workload].get("slab"):
        args.warmup = max(args.warmup, 3)  # the 8192x8192 slab workload takes about a second per step: 1 warm-up is allowed
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import flow2d_loader
    flow2d_loader.load()
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload], rank, world)
    else:
        run_ours(args, args.workload, rank, world, local_rank)


if __name__ == "__main__":
    main()
