"""A/B of the thread-block-cluster solve (csrc/solve_cluster.cu) on one GPU, one process: the switches are read when a
handle is created, so every configuration is measured back to back on the same box.

  python tools/ab_cluster.py [out.json]

1. per level of the 1024^2 pyramid (BASELINE configs[3]: 40 x 5 iterations): time of one level's solve through
   flow2d_stage_solve, CUDA events on the launching stream, for every configuration -> where the cluster kernel wins and
   what a barrier-to-barrier phase costs (the scheduler's time model, flow2d_api.cu: cluster_phase_us)
2. the workloads of bench.py (C4 batch = the default bench line, C4 one pair at a time, C1b) with bench.measure
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    "off": dict(FLOW2D_CLUSTER="0"),
    "default": dict(),  # by the number of handles alive on the device: "whole" from 4 on, "off" below
    "small": dict(FLOW2D_CLUSTER="1"),  # the cluster where a CTA's block fits 256 threads (levels of 1 025 .. 4 096 px)
    "whole": dict(FLOW2D_CLUSTER="2"),  # every level that fits a cluster (.. 16 384 px)
    "whole_compact": dict(FLOW2D_CLUSTER="2", FLOW2D_CLUSTER_COMPACT="1"),
    "whole_max8": dict(FLOW2D_CLUSTER="2", FLOW2D_CLUSTER_MAX="8"),
    "whole_pass": dict(FLOW2D_CLUSTER="2", FLOW2D_CLUSTER_PASS="1"),
    "whole_pass_forced": dict(FLOW2D_CLUSTER="2", FLOW2D_CLUSTER_PASS="2"),
}
QUICK = ("off", "default", "whole", "off")  # `python tools/ab_cluster.py out.json quick`: the workloads only, these configurations
KEYS = ("FLOW2D_CLUSTER", "FLOW2D_CLUSTER_PASS", "FLOW2D_CLUSTER_COMPACT", "FLOW2D_CLUSTER_MAX")


def set_env(cfg):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(cfg)


def level_times(torch, m, names):
    import bench
    wl = bench.WORKLOADS["c4"]
    w, h, cfg = wl["w"], wl["h"], wl["cfg"]
    f0, f1 = bench.make_frames(wl, 0)
    table = m.level_table(w, h, cfg["scale"], cfg["levels"])
    rows = {}
    for name in names:
        set_env(CONFIGS[name])
        fl = m.Flow2D(w, h)
        st = torch.cuda.Stream()
        fl.set_stream(st.cuda_stream)
        d0, d1 = fl.to_container(f0), fl.to_container(f1)
        t = [fl.container(0.0) for _ in range(4)]
        sp = m.default_params(**cfg)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for cw, ch, hx, hy in table:
            if cw * ch <= 1024 or cw * ch > 300000:
                continue
            c0 = fl.launch_counts()
            with torch.cuda.stream(st):
                fl.stage_solve(d0, d1, t[0], t[1], t[2], t[3], None, None, cw, ch, float(hx), float(hy), sp)
                c1 = fl.launch_counts()
                a.record(st)
                for _ in range(3):
                    fl.stage_solve(d0, d1, t[0], t[1], t[2], t[3], None, None, cw, ch, float(hx), float(hy), sp)
                b.record(st)
            torch.cuda.synchronize()
            used = {k: v - c0.get(k, 0) for k, v in c1.items() if k.startswith("solve") and v > c0.get(k, 0)}
            rows.setdefault("%dx%d" % (cw, ch), {})[name] = {"us": round(a.elapsed_time(b) / 3 * 1e3, 1), "kernels": used}
        fl.destroy()
    return rows


def main():
    import torch
    import flow2d_loader
    m = flow2d_loader.load()
    import bench
    quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
    out = {"levels_c4": level_times(torch, m, ["off", "default"] if quick else list(CONFIGS))}
    for lv, r in out["levels_c4"].items():
        print(lv, {k: v["us"] for k, v in r.items()}, file=sys.stderr)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(json.dumps(out, indent=1))
    ctx = bench.Ctx()
    ctx.torch, ctx.m, ctx.rank, ctx.world, ctx.dev, ctx.dist, ctx.flush = torch, m, 0, 1, 0, None, None
    torch.cuda.set_device(0)
    runs = [("c4_batch", "c4", 6, 3, 0, 0), ("c4_single", "c4", 10, 3, 1, 1), ("c1b_single", "c1b", 20, 3, 1, 1), ("c1b_batch", "c1b", 6, 3, 8, 16)]
    out["workloads"] = {}
    for label, key, steps, warmup, streams, pairs in runs:
        for name in (QUICK if quick else ("off", "default", "whole", "whole_compact", "whole_pass", "off")):
            set_env(CONFIGS[name])
            try:
                res, _ = bench.measure(ctx, key, steps, warmup, streams, pairs, detail=False)
                rec = {"value": round(res["value"], 2), "e2e": round(res["e2e"]["value"], 2), "ms_per_step": round(res["ms_per_step"], 3),
                       "launches": res["launches_by_kernel"]}
            except Exception as e:  # keep what has been measured
                rec = {"error": repr(e)}
            out["workloads"].setdefault(label, {}).setdefault(name, []).append(rec)
            print(label, name, rec, file=sys.stderr)
            if len(sys.argv) > 1:
                open(sys.argv[1], "w").write(json.dumps(out, indent=1))
    text = json.dumps(out, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
