"""Runs the reference's own CUDA build (oracle/_ref) on the GPU box and stores its outputs as
golden vectors.  Usage (under gpurun):  python tools/make_golden.py gpurun_out/golden
The resulting .npz files are committed under tests/golden/ so that the CPU-only test-suite can pin
the oracle against the reference without a GPU."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import RefHarness, ref_available  # noqa: E402

C1B = dict(levels=50, scale=0.9, outer=40, inner=5, alpha=35.0, e_smooth=0.001, e_data=0.001, median=5, sigma=1.5)
C1A = dict(levels=20, scale=0.9, outer=20, inner=5, alpha=3.5, e_smooth=0.001, e_data=0.001, median=5, sigma=0.45)


def main(out):
    os.makedirs(out, exist_ok=True)
    z = np.load(os.path.join(ROOT, "tests", "golden", "rub_u8.npz"))
    f0, f1 = z["rub1"].astype(np.float32), z["rub2"].astype(np.float32)
    for variant in ("", "cubin"):
        if not ref_available(variant):
            print("reference build variant %r missing" % variant)
            continue
        tag = variant or "ptx"
        with tempfile.TemporaryDirectory() as tmp:
            ref = RefHarness(tmp, variant)
            for name, cfg in (("c1b", C1B), ("c1a", C1A)):
                u, v, ms = ref.flow(f0, f1, cfg, warmup=1, reps=3)
                print(tag, name, "REF_MS", ms, "u", float(u.min()), float(u.max()), "v", float(v.min()), float(v.max()))
                np.savez_compressed(os.path.join(out, "rub_%s_reference_%s.npz" % (name, tag)), u=u, v=v, ms=np.array(ms))
            # small per-stage goldens for the CPU suite
            import flow2d_loader
            S = __import__("importlib").import_module("cuda_flow2d_b200.synth") if flow2d_loader.load() else None
            w, h, hx, hy = 53, 41, 2.92, 2.425
            a0, a1, _, _ = S.make_pair(w, h, 300, U1=1.5, L=48.0)
            u0 = S.smooth_random(w, h, 11, -2, 2)
            v0 = S.smooth_random(w, h, 12, -2, 2)
            du, dv, phi, ksi = ref.solve(a0, a1, u0, v0, hx, hy, 20.0, 0.001, 0.001, 3, 5, 0)
            gdu, gdv, _, _ = ref.solve(a0, a1, u0, v0, hx, hy, 20.0, 0.001, 0.001, 2, 5, 1)
            rng = np.random.default_rng(9)
            img = rng.uniform(0, 255, (h, w)).astype(np.float32)
            stage = dict(f0=a0, f1=a1, u=u0, v=v0, hx=hx, hy=hy, du=du, dv=dv, phi=phi, ksi=ksi, grad_du=gdu, grad_dv=gdv,
                         img=img, blur=ref.conv(img, 1.5), down=ref.resample(img, 23, 19), up=ref.resample(img, 61, 47),
                         warp=ref.warp(a0, a1, (u0 * hx).astype(np.float32), (v0 * hy).astype(np.float32), hx, hy),
                         med5=ref.median(u0, 5), med3=ref.median(u0, 3), med7=ref.median(u0, 7))
            np.savez_compressed(os.path.join(out, "stages_reference_%s.npz" % tag), **stage)
            print(tag, "stage goldens written")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
