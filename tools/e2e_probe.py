"""Where does the end-to-end time of the C4 batch go?  Host time of the enqueue calls, a step synchronised like bench.py
(every handle after every step) against the same pairs pipelined without step boundaries.
    python tools/e2e_probe.py [handles] [pairs]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import flow2d_loader  # noqa: E402

m = flow2d_loader.load()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P = int(sys.argv[2]) if len(sys.argv) > 2 else 16
wl = bench.WORKLOADS["c4"]
w, h = wl["w"], wl["h"]
frames = [bench.make_frames(wl, i) for i in range(4)]
handles = [m.Flow2D(w, h) for _ in range(K)]
streams = [torch.cuda.Stream() for _ in range(K)]
for hd, st in zip(handles, streams):
    hd.set_stream(st.cuda_stream)
p = m.default_params(**wl["cfg"])
hin = [(torch.from_numpy(frames[i % 4][0]).pin_memory(), torch.from_numpy(frames[i % 4][1]).pin_memory()) for i in range(P)]
hout = [(torch.empty((h, w)).pin_memory(), torch.empty((h, w)).pin_memory()) for _ in range(P)]


def step(sync=True):
    t0 = time.perf_counter()
    for i in range(P):
        handles[i % K].compute_async(hin[i][0], hin[i][1], p, hout[i][0], hout[i][1])
    t1 = time.perf_counter()
    if sync:
        for hd in handles:
            hd.synchronize()
    return t1 - t0


for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
enq = [step() for _ in range(4)]
torch.cuda.synchronize()
t_sync = (time.perf_counter() - t0) / 4
t0 = time.perf_counter()
enq2 = [step(sync=False) for _ in range(4)]
for hd in handles:
    hd.synchronize()
t_pipe = (time.perf_counter() - t0) / 4
print("handles %d pairs %d: host enqueue %.2f ms per step (%.3f ms per call); step with per-step sync %.2f ms = %.1f Mpix/s; "
      "pipelined %.2f ms per step = %.1f Mpix/s (enqueue %.2f ms)" % (K, P, 1e3 * sum(enq) / 4, 1e3 * sum(enq) / 4 / P, 1e3 * t_sync,
                                                                       P * w * h / t_sync / 1e6, 1e3 * t_pipe, P * w * h / t_pipe / 1e6,
                                                                       1e3 * sum(enq2) / 4))
