"""End-to-end throughput of the sequence driver (cuda-flow2d --sequence, host/sequence.cpp): C4-like frames
(1024x1024, low contrast, noisy) are written as ONE 8-bit stack and as float32 files, the command
line computes the flow of every consecutive pair and writes flow-u / flow-v files.  The timed region is the
driver's own clock: first frame read -> last output file closed (file I/O, H2D, compute, D2H).
    python tools/bench_sequence.py [pairs=32] [handles per GPU=8] [dir=/dev/shm] [devices, e.g. 0-7]
With several devices pair i runs on GPU i mod N (cuda-flow2d --sequence --devices ...): the product form of BASELINE.json
configs[3] ("batch of frame pairs sharded one pair per GPU")."""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402
import flow2d_loader  # noqa: E402

flow2d_loader.load()

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
handles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
where = sys.argv[3] if len(sys.argv) > 3 else ("/dev/shm" if os.path.isdir("/dev/shm") else None)
devices = sys.argv[4] if len(sys.argv) > 4 else None
wl = bench.WORKLOADS["c4"]
w, h = wl["w"], wl["h"]
cli = os.path.join(ROOT, "cuda-flow2d_b200", "bin", "cuda-flow2d")
with tempfile.TemporaryDirectory(dir=where) as tmp:
    # 8 distinct frames cycled: consecutive frames differ, the generator stays off the clock
    distinct = [bench.make_frames(wl, i)[0] for i in range(8)]
    frames = [distinct[i % 8] for i in range(pairs + 1)]
    os.mkdir(os.path.join(tmp, "out"))
    os.mkdir(os.path.join(tmp, "f32"))
    np.stack([np.clip(np.rint(f), 0, 255).astype(np.uint8) for f in frames]).tofile(os.path.join(tmp, "stack_u8.raw"))
    for i, f in enumerate(frames):
        f.tofile(os.path.join(tmp, "f32", "frame%04d.raw" % i))
    res = {}
    for name, src in (("stack_u8", "stack_u8.raw"), ("dir_f32", "f32")):
        best = None
        for rep in range(3):
            r = subprocess.run([cli, "--sequence", str(w), str(h), "out/", src, "--handles", str(handles)] +
                               (["--devices", devices] if devices else []), cwd=tmp,
                               stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
            out = r.stdout.decode()
            m = re.search(r"in ([0-9.]+) s: ([0-9.]+) pairs/s, ([0-9.]+) Mpix/s", out)
            if r.returncode != 0 or not m:
                print(out[-2000:])
                sys.exit(1)
            busy = re.search(r"reader busy ([0-9.]+) s, writer busy ([0-9.]+) s; scheduler waited ([0-9.]+) s for frames, "
                             r"([0-9.]+) s for the GPU, ([0-9.]+) s for the writer", out)
            cur = {"seconds": float(m.group(1)), "pairs_per_s": float(m.group(2)), "mpix_per_s": float(m.group(3)),
                   "reader_busy_s": float(busy.group(1)), "writer_busy_s": float(busy.group(2)),
                   "wait_frames_s": float(busy.group(3)), "wait_gpu_s": float(busy.group(4)), "wait_writer_s": float(busy.group(5))}
            if best is None or cur["mpix_per_s"] > best["mpix_per_s"]:
                best = cur
        res[name] = best
print(json.dumps({"metric": "sequence driver, end to end incl. file I/O", "unit": "Mpix/s", "pairs": pairs, "handles_per_gpu": handles,
                  "devices": devices or "0",
                  "frame": [w, h], "settings": wl["cfg"], "storage": where or "tmp", "best_of": 3, "inputs": res,
                  "bytes_per_pair": {"read_u8": w * h, "read_f32": 4 * w * h, "written": 8 * w * h}}))
