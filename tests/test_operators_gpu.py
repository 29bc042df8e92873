"""The source-compatible operator classes (cuda-flow2d_b200/host/cuda_operations.h: CudaOperation*2D with the reference's
Initialize / Execute / Destroy and by-name parameters, src/cuda_operations/cuda_operation_base.h:44-51) driven by a C++
program written the way OpticalFlow2D::ComputeFlow drives the reference's operators; every result against the oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "cuda-flow2d_b200", "bin", "operators_test")


def test_operator_classes_against_the_oracle(pkg, oracle, synth, tmp_path):
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    w, h = 200, 150
    f0, f1, _, _ = synth.make_pair(w, h, 21, U1=1.5)
    u = synth.smooth_random(w, h, 1, -1, 1)
    v = synth.smooth_random(w, h, 2, -1, 1)
    for name, a in (("f0", f0), ("f1", f1), ("u", u), ("v", v)):
        a.astype(np.float32).tofile(tmp_path / (name + ".raw"))
    r = subprocess.run([EXE, str(w), str(h), str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    out = r.stdout.decode()
    assert r.returncode == 0 and "OPERATORS_DONE" in out, out[-2000:]
    # the reference's messages for a missing key and for in-place use
    assert "Missing parameter 'dev_output'" in out and "Input buffer cannot serve as output buffer" in out

    def rd(name, ww=w, hh=h):
        return np.fromfile(tmp_path / (name + ".raw"), np.float32).reshape(hh, ww)

    hx, hy = np.float32(1.25), np.float32(1.5)
    assert np.array_equal(rd("blur"), oracle.blur(f0, 1.5))
    rw, rh = (w * 7 + 9) // 10, (h * 7 + 9) // 10
    assert np.array_equal(rd("resample", rw, rh), oracle.resample(f0, rw, rh))
    warped = oracle.warp(f0, f1, u, v, hx, hy)
    assert np.array_equal(rd("warp"), warped, equal_nan=True)
    du, dv, phi, ksi = oracle.solve_level(f0, warped, u, v, hx, hy, oracle.make_params(outer=3, inner=5, alpha=20.0))
    assert np.array_equal(rd("du"), du) and np.array_equal(rd("dv"), dv)
    assert np.array_equal(rd("phi"), phi) and np.array_equal(rd("ksi"), ksi)
    added = u + du
    assert np.array_equal(rd("add"), added)
    med, rc = oracle.median(added, 5)
    assert rc == 0 and np.array_equal(rd("median"), med)
