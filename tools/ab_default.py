"""The shipped default of the cluster solve, measured: no switch set, C4 batch (8 handles: cluster schedule) and one C4 /
C1b pair at a time (lone handle: no cluster).  python tools/ab_default.py out.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for k in ("FLOW2D_CLUSTER", "FLOW2D_CLUSTER_PASS", "FLOW2D_CLUSTER_COMPACT", "FLOW2D_CLUSTER_MAX"):
    os.environ.pop(k, None)
import torch  # noqa: E402

import flow2d_loader  # noqa: E402

m = flow2d_loader.load()
import bench  # noqa: E402

ctx = bench.Ctx()
ctx.torch, ctx.m, ctx.rank, ctx.world, ctx.dev, ctx.dist, ctx.flush = torch, m, 0, 1, 0, None, None
torch.cuda.set_device(0)
out = {}
for label, key, steps, warmup, streams, pairs in (("c4_batch", "c4", 6, 3, 0, 0), ("c4_single", "c4", 10, 3, 1, 1), ("c1b_single", "c1b", 20, 3, 1, 1)):
    res, _ = bench.measure(ctx, key, steps, warmup, streams, pairs, detail=False)
    out[label] = {"value": round(res["value"], 2), "e2e": round(res["e2e"]["value"], 2), "ms_per_step": round(res["ms_per_step"], 3),
                  "launches": res["launches_by_kernel"]}
    print(label, out[label], file=sys.stderr)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(json.dumps(out, indent=1))
