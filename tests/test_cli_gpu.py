"""The cuda-flow2d command line (cuda-flow2d_b200/bin/cuda-flow2d) against the reference's own
executable (oracle/_ref/cuda-flow2d, its unmodified main.cpp) on the bundled rub pair: same argv
form, and byte-identical flow-u / flow-v / amp / res.pgm files."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REF_DIR, ROOT

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "cuda-flow2d_b200", "bin", "cuda-flow2d")
REF_CLI = os.path.join(REF_DIR, "cuda-flow2d")


def _run(exe, args, cwd):
    return subprocess.run([exe] + [str(a) for a in args], cwd=cwd, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, timeout=600)


def test_argv_form_matches_reference_files(rub, tmp_path):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/cuda-flow2d is not present")
    f0, f1 = rub
    (tmp_path / "ours").mkdir()
    (tmp_path / "ref").mkdir()
    f0.tofile(tmp_path / "rub1_f32.raw")
    f1.tofile(tmp_path / "rub2_f32.raw")
    # cuda-flow2d <file1> <file2> <W> <H> <counter> <output path>   (src/main.cpp:107-118, defaults 70-80)
    r = _run(CLI, ["rub1_f32.raw", "rub2_f32.raw", 584, 388, "t_", "ours/"], tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    assert b"Total GPU computation time" in r.stdout
    q = _run(REF_CLI, ["rub1_f32.raw", "rub2_f32.raw", 584, 388, "t_", "ref/"], tmp_path)
    assert q.returncode == 0, q.stdout.decode()[-2000:]
    for name in ("t_flow-u-584-388.raw", "t_flow-v-584-388.raw", "t_amp-584-388.raw", "t_res.pgm"):
        a = (tmp_path / "ours" / name).read_bytes()
        b = (tmp_path / "ref" / name).read_bytes()
        assert len(a) == len(b) and a == b, "%s differs (%d vs %d bytes)" % (name, len(a), len(b))


def test_settings_form_8bit_and_exit_codes(rub, oracle, tmp_path):
    f0, f1 = rub
    f0.astype(np.uint8).tofile(tmp_path / "rub1.raw")  # the bundled files are 8-bit
    f1.astype(np.uint8).tofile(tmp_path / "rub2.raw")
    (tmp_path / "out").mkdir()
    xml = """<?xml version="1.0"?>
<!-- settings.xml schema of the reference -->
<OpticalFlow>
  <Input><Path inputPath="%s/"/><Mode Nx="584" Ny="388" imageType="8-bit"><Files file1 ="rub1.raw" file2 ="rub2.raw"/></Mode></Input>
  <Parameters><Method mode ="2d" run="flow" key="0" />
    <Solver><Iterations inner="5" outer="20"/><Warping levels="20" scaling="0.9" medianRadius="5"/>
      <Model sigma="0.45" alpha ="3.5" e_smooth="0.001" e_data="0.001"/></Solver></Parameters>
  <Output><Path outputPath="out/"/></Output>
</OpticalFlow>""" % tmp_path
    (tmp_path / "settings.xml").write_text(xml)
    r = _run(CLI, [], tmp_path)  # no argument: settings.xml in the current directory
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    u = np.fromfile(tmp_path / "out" / "flow-u-584-388.raw", np.float32).reshape(388, 584)
    v = np.fromfile(tmp_path / "out" / "flow-v-584-388.raw", np.float32).reshape(388, 584)
    ou, ov = oracle.compute_flow(f0, f1, oracle.make_params(levels=20, outer=20, alpha=3.5, sigma=0.45))
    assert np.all(u == ou) and np.all(v == ov)
    z = np.load(os.path.join(ROOT, "tests", "golden", "rub_c1a_reference.npz"))
    assert np.all(u == z["u"]) and np.all(v == z["v"])  # = the reference build's flow for these settings
    # exit codes of src/main.cpp: 3 = settings unreadable, 2 = input unreadable, 0 = usage
    assert _run(CLI, ["missing.xml"], tmp_path).returncode == 3
    assert _run(CLI, ["nope1.raw", "nope2.raw", 584, 388, "x_", "out/"], tmp_path).returncode == 2
    assert _run(CLI, ["a", "b", "c"], tmp_path).returncode == 0


def test_sequence_mode_equals_pairwise(synth, oracle, tmp_path):
    """--sequence: flows between consecutive frames, several pairs in flight on concurrent handles;
    every pair must equal the single-pair result (= oracle) bit for bit."""
    w, h, n = 96, 80, 7
    (tmp_path / "out").mkdir()
    frames = []
    for i in range(n):
        f, _, _, _ = synth.make_pair(w, h, 100, U0=(0.4 * i, -0.2 * i), U1=0.0)  # the texture translating steadily
        frames.append(f)
        f.tofile(tmp_path / ("frame%02d.raw" % i))
    names = ["frame%02d.raw" % i for i in range(n)]
    r = _run(CLI, ["--sequence", w, h, "out/"] + names, tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    p = oracle.make_params()
    for i in range(n - 1):
        u = np.fromfile(tmp_path / "out" / ("%04d_flow-u-%d-%d.raw" % (i, w, h)), np.float32).reshape(h, w)
        v = np.fromfile(tmp_path / "out" / ("%04d_flow-v-%d-%d.raw" % (i, w, h)), np.float32).reshape(h, w)
        ou, ov = oracle.compute_flow(frames[i], frames[i + 1], p)
        assert np.all(u == ou) and np.all(v == ov), "pair %d" % i


def test_sequence_stack_u8_directory_and_writers(rub, oracle, tmp_path):
    """--sequence on an 8-bit multi-frame stack and on a directory of float32 frames (frame source auto-detection,
    reader / writer threads, 2K output slots): every pair equals the oracle bit for bit; --color / --amp write the
    same res.pgm / amp files as the single-pair command line does for that pair."""
    w, h, n = 72, 56, 11
    rng = np.random.default_rng(5)
    base = rng.integers(0, 255, (h + 2 * n, w + 2 * n), dtype=np.uint8)
    k = np.ones(5) / 5.0
    smooth = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 1, base.astype(np.float64))
    smooth = np.apply_along_axis(lambda c: np.convolve(c, k, "same"), 0, smooth)
    frames = [np.ascontiguousarray(smooth[i:i + h, n - i:n - i + w]).astype(np.uint8) for i in range(n)]  # a moving window
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    (tmp_path / "one").mkdir()
    (tmp_path / "dir").mkdir()
    np.stack(frames).tofile(tmp_path / "stack.raw")
    for i, f in enumerate(frames):
        f.astype(np.float32).tofile(tmp_path / "dir" / ("frame%03d.raw" % i))
    r = _run(CLI, ["--sequence", w, h, "a/", "stack.raw", "--handles", 3, "--color", "--amp"], tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    assert b"11 frames of 72x56 (stack, 8-bit)" in r.stdout and b"10 frame pairs on 3 concurrent handles" in r.stdout
    r = _run(CLI, ["--sequence", w, h, "b/", "dir", "--handles", 8], tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    assert b"(directory, float32)" in r.stdout
    p = oracle.make_params()
    for i in range(n - 1):
        ou, ov = oracle.compute_flow(frames[i].astype(np.float32), frames[i + 1].astype(np.float32), p)
        for d in ("a", "b"):
            u = np.fromfile(tmp_path / d / ("%04d_flow-u-%d-%d.raw" % (i, w, h)), np.float32).reshape(h, w)
            v = np.fromfile(tmp_path / d / ("%04d_flow-v-%d-%d.raw" % (i, w, h)), np.float32).reshape(h, w)
            assert np.all(u == ou) and np.all(v == ov), "pair %d (%s)" % (i, d)
    # writers: pair 4 through the single-pair command line
    r = _run(CLI, ["dir/frame004.raw", "dir/frame005.raw", w, h, "p4_", "one/"], tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    for seq_name, one_name in (("0004_res.pgm", "p4_res.pgm"), ("0004_amp-%d-%d.raw" % (w, h), "p4_amp-%d-%d.raw" % (w, h))):
        assert (tmp_path / "a" / seq_name).read_bytes() == (tmp_path / "one" / one_name).read_bytes(), seq_name


def test_sequence_gradient_settings_and_read_failure(synth, oracle, tmp_path):
    """--settings takes the solver values of a settings.xml, --gradient the data term; a frame that disappears
    mid-sequence ends the run with exit code 2 without hanging the pipeline."""
    w, h, n = 64, 48, 5
    (tmp_path / "out").mkdir()
    frames = []
    for i in range(n):
        f, _, _, _ = synth.make_pair(w, h, 7, U0=(0.3 * i, 0.1 * i), U1=0.0)
        frames.append(f)
        f.tofile(tmp_path / ("s%02d.raw" % i))
    (tmp_path / "s.xml").write_text(
        '<Settings><Input><Path inputPath="./"/><Mode Nx="1" Ny="1" imageType="32-bit"><Files file1="x" file2="y"/></Mode></Input>'
        '<Parameters><Method key="false"/><Solver><Iterations inner="3" outer="4"/>'
        '<Warping levels="6" scaling="0.8" medianRadius="3"/><Model sigma="0.8" alpha="12" e_smooth="0.01" e_data="0.02"/></Solver>'
        '</Parameters>'
        '<Output><Path outputPath="./"/></Output></Settings>')
    names = ["s%02d.raw" % i for i in range(n)]
    r = _run(CLI, ["--sequence", w, h, "out/", "--settings", "s.xml", "--gradient"] + names, tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    p = oracle.make_params(levels=6, scale=0.8, outer=4, inner=3, alpha=12.0, e_smooth=0.01, e_data=0.02, median=3, sigma=0.8,
                           constancy=oracle.GRADIENT)
    for i in range(n - 1):
        u = np.fromfile(tmp_path / "out" / ("%04d_flow-u-%d-%d.raw" % (i, w, h)), np.float32).reshape(h, w)
        v = np.fromfile(tmp_path / "out" / ("%04d_flow-v-%d-%d.raw" % (i, w, h)), np.float32).reshape(h, w)
        ou, ov = oracle.compute_flow(frames[i], frames[i + 1], p)
        assert np.all(u == ou) and np.all(v == ov), "pair %d" % i
    # a truncated frame in the middle of a stack of files is caught by the size check up front ...
    (tmp_path / "s02.raw").write_bytes(b"1234")
    assert _run(CLI, ["--sequence", w, h, "out/"] + names, tmp_path).returncode == 2
    # ... an output directory that does not exist by the writer thread (exit code 4)
    frames[2].tofile(tmp_path / "s02.raw")
    assert _run(CLI, ["--sequence", w, h, "missing_dir/"] + names, tmp_path).returncode == 4


def test_pair_form_with_residuals_flag(rub, tmp_path):
    """Trailing --residuals (not in the reference): same output files, plus one residual line per pyramid level."""
    f0, f1 = rub
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    f0.tofile(tmp_path / "r1.raw")
    f1.tofile(tmp_path / "r2.raw")
    r = _run(CLI, ["r1.raw", "r2.raw", 584, 388, "x_", "a/"], tmp_path)
    q = _run(CLI, ["r1.raw", "r2.raw", 584, 388, "x_", "b/", "--residuals"], tmp_path)
    assert r.returncode == 0 and q.returncode == 0, q.stdout.decode()[-2000:]
    lines = [l for l in q.stdout.decode().splitlines() if l.startswith("Level ") and "residual rms" in l]
    assert len(lines) == 47 and b"residual rms" not in r.stdout
    vals = [float(l.split()[-3]) for l in lines] + [float(l.split()[-1]) for l in lines]
    assert all(np.isfinite(v) and v >= 0 for v in vals)
    for name in ("x_flow-u-584-388.raw", "x_flow-v-584-388.raw", "x_amp-584-388.raw", "x_res.pgm"):
        assert (tmp_path / "a" / name).read_bytes() == (tmp_path / "b" / name).read_bytes(), name


def test_settings_form_float32_frames_with_stock_image_type(rub, tmp_path):
    """The reference ignores Mode@imageType and always reads float32 (src/main.cpp:175-176), while its stock settings.xml
    says imageType="8-bit": float32 frames with an unchanged template must load (the reader follows the file size)."""
    f0, f1 = rub
    f0.tofile(tmp_path / "a.raw")
    f1.tofile(tmp_path / "b.raw")
    (tmp_path / "out").mkdir()
    xml = """<OpticalFlow>
  <Input><Path inputPath="./"/><Mode Nx="584" Ny="388" imageType="8-bit"><Files file1="a.raw" file2="b.raw"/></Mode></Input>
  <Parameters><Method mode="2d" run="flow" key="0"/>
    <Solver><Iterations inner="5" outer="20"/><Warping levels="20" scaling="0.9" medianRadius="5"/>
      <Model sigma="0.45" alpha="3.5" e_smooth="0.001" e_data="0.001"/></Solver></Parameters>
  <Output><Path outputPath="out/"/></Output></OpticalFlow>"""
    (tmp_path / "settings.xml").write_text(xml)
    r = _run(CLI, [], tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    u = np.fromfile(tmp_path / "out" / "flow-u-584-388.raw", np.float32).reshape(388, 584)
    z = np.load(os.path.join(ROOT, "tests", "golden", "rub_c1a_reference.npz"))
    assert np.all(u == z["u"])


def test_failed_solve_is_not_reported_as_success(rub, tmp_path):
    """A parameter the solver refuses (median size 9): non-zero exit code and no output files (upstream writes stale memory)."""
    f0, f1 = rub
    f0.tofile(tmp_path / "a.raw")
    f1.tofile(tmp_path / "b.raw")
    (tmp_path / "out").mkdir()
    xml = """<OpticalFlow>
  <Input><Path inputPath="./"/><Mode Nx="584" Ny="388" imageType="32-bit"><Files file1="a.raw" file2="b.raw"/></Mode></Input>
  <Parameters><Method mode="2d" run="flow" key="0"/>
    <Solver><Iterations inner="5" outer="2"/><Warping levels="3" scaling="0.9" medianRadius="9"/>
      <Model sigma="0.45" alpha="3.5" e_smooth="0.001" e_data="0.001"/></Solver></Parameters>
  <Output><Path outputPath="out/"/></Output></OpticalFlow>"""
    (tmp_path / "settings.xml").write_text(xml)
    r = _run(CLI, [], tmp_path)
    assert r.returncode not in (0, 1, 2, 3), r.stdout.decode()[-2000:]
    assert not (tmp_path / "out" / "flow-u-584-388.raw").exists()


def test_sequence_on_several_devices(synth, oracle, tmp_path):
    """--sequence --devices: pair i -> device list entry i mod N, K handles per entry, several reader / writer threads.
    On a one-GPU box the list names device 0 twice (same scheduling code, two handle groups); with more GPUs the
    driver's own multi-GPU run (tools/bench_sequence.py ... 0-7) covers the real thing.  Every pair equals the oracle."""
    import torch
    w, h, n = 96, 80, 9
    (tmp_path / "out").mkdir()
    frames = []
    for i in range(n):
        f, _, _, _ = synth.make_pair(w, h, 100, U0=(0.4 * i, -0.2 * i), U1=0.0)
        frames.append(f)
        f.tofile(tmp_path / ("frame%02d.raw" % i))
    names = ["frame%02d.raw" % i for i in range(n)]
    devs = "0-%d" % (min(torch.cuda.device_count(), 4) - 1) if torch.cuda.device_count() > 1 else "0,0"
    r = _run(CLI, ["--sequence", w, h, "out/"] + names + ["--devices", devs, "--handles", 2, "--io-threads", 2], tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    assert b"2 reader and 2 writer" in r.stdout
    p = oracle.make_params()
    for i in range(n - 1):
        u = np.fromfile(tmp_path / "out" / ("%04d_flow-u-%d-%d.raw" % (i, w, h)), np.float32).reshape(h, w)
        v = np.fromfile(tmp_path / "out" / ("%04d_flow-v-%d-%d.raw" % (i, w, h)), np.float32).reshape(h, w)
        ou, ov = oracle.compute_flow(frames[i], frames[i + 1], p)
        assert np.all(u == ou) and np.all(v == ov), "pair %d" % i


def test_slab_mode_command_line(synth, tmp_path):
    """cuda-flow2d --slab: one frame pair on several GPUs (the same device twice on a one-GPU box: two ranks, real
    mailboxes and flags), output files byte-identical to the single-GPU command line."""
    import torch
    w, h = 160, 720
    f0, f1, _, _ = synth.make_pair(w, h, 9, U0=(0.5, -0.3), U1=0.8, L=64.0)
    f0.tofile(tmp_path / "a.raw")
    f1.tofile(tmp_path / "b.raw")
    (tmp_path / "one").mkdir()
    (tmp_path / "two").mkdir()
    r = _run(CLI, ["a.raw", "b.raw", w, h, "one/"], tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    devs = "0-1" if torch.cuda.device_count() > 1 else "0,0"
    r = _run(CLI, ["--slab", devs, "a.raw", "b.raw", w, h, "two/"], tmp_path)
    assert r.returncode == 0, r.stdout.decode()[-2000:]
    for name in ("flow-u-%d-%d.raw" % (w, h), "flow-v-%d-%d.raw" % (w, h), "res.pgm", "amp-%d-%d.raw" % (w, h)):
        assert (tmp_path / "one" / name).read_bytes() == (tmp_path / "two" / name).read_bytes(), name
