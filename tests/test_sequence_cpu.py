"""Host logic of the sequence driver (cuda-flow2d_b200/host/sequence.cpp) that needs no GPU:
where frames come from (file list / directory / multi-frame stack) and the 8-bit / float32 detection by size
(the reference has one reader per type, src/data_types/data2d.cpp:98-141 and :143-178)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CLI = os.path.join(ROOT, "cuda-flow2d_b200", "bin", "cuda-flow2d")
W, H = 40, 30


def _list(args, cwd):
    r = subprocess.run([CLI, "--sequence", str(W), str(H), "out/"] + [str(a) for a in args] + ["--list"], cwd=cwd,
                       stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=60)
    return r.returncode, r.stdout.decode()


@pytest.fixture()
def frames(tmp_path):
    rng = np.random.default_rng(3)
    fr = [rng.integers(0, 255, (H, W), dtype=np.uint8) for _ in range(8)]
    (tmp_path / "dir8").mkdir()
    (tmp_path / "dir32").mkdir()
    for i, f in enumerate(fr):
        f.tofile(tmp_path / "dir8" / ("f%02d.raw" % (7 - i)))          # names sort in reverse creation order
        f.astype(np.float32).tofile(tmp_path / "dir32" / ("g%02d.raw" % i))
    np.stack(fr[:5]).tofile(tmp_path / "stack5_u8.raw")
    np.stack(fr).tofile(tmp_path / "stack8_u8.raw")                    # as long as 2 float32 frames
    np.stack(fr[:3]).astype(np.float32).tofile(tmp_path / "stack3_f32.raw")
    return tmp_path


def test_directory_is_sorted_by_name_and_typed_by_size(frames):
    rc, out = _list(["dir8"], frames)
    assert rc == 0 and "8 frames of 40x30 (directory, 8-bit)" in out
    names = [l.split()[1] for l in out.splitlines() if l.startswith("  0")]
    assert names == ["dir8/f%02d.raw" % i for i in range(8)]
    rc, out = _list(["dir32"], frames)
    assert rc == 0 and "8 frames of 40x30 (directory, float32)" in out


def test_file_list_keeps_the_given_order(frames):
    rc, out = _list(["dir32/g03.raw", "dir32/g01.raw", "dir32/g02.raw"], frames)
    assert rc == 0 and "3 frames of 40x30 (files, float32)" in out
    assert [l.split()[1] for l in out.splitlines() if l.startswith("  0")] == ["dir32/g03.raw", "dir32/g01.raw", "dir32/g02.raw"]


def test_stack_detection_and_override(frames):
    rc, out = _list(["stack5_u8.raw"], frames)
    assert rc == 0 and "5 frames of 40x30 (stack, 8-bit)" in out
    rc, out = _list(["stack3_f32.raw"], frames)
    assert rc == 0 and "3 frames of 40x30 (stack, float32)" in out
    # 8 x u8 == 2 x f32 bytes: float32 wins the tie, --u8 overrides
    rc, out = _list(["stack8_u8.raw"], frames)
    assert rc == 0 and "2 frames of 40x30 (stack, float32)" in out
    rc, out = _list(["stack8_u8.raw", "--u8"], frames)
    assert rc == 0 and "8 frames of 40x30 (stack, 8-bit)" in out


def test_bad_inputs_exit_2_like_unreadable_frames(frames):
    """Exit code 2 = input unreadable, as for the reference's pair form (src/main.cpp:178-182)."""
    assert _list(["dir8/f00.raw", "dir32/g00.raw"], frames)[0] == 2      # mixed pixel types
    assert _list(["nope.raw", "nope2.raw"], frames)[0] == 2
    assert _list(["dir32/g00.raw"], frames)[0] == 2                      # one frame is not a sequence
    (frames / "odd.raw").write_bytes(b"x" * (W * H + 1))
    assert _list(["dir8/f00.raw", "odd.raw"], frames)[0] == 2
    assert _list(["dir32", "--u8"], frames)[0] == 2                      # forced type does not match the sizes
    (frames / "empty").mkdir()
    assert _list(["empty"], frames)[0] == 2


def _cli(args, cwd):
    r = subprocess.run([CLI] + [str(a) for a in args], cwd=cwd, stdin=subprocess.DEVNULL, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=60)
    return r.returncode, r.stdout.decode()


def test_cli_exit_codes_that_need_no_gpu(tmp_path):
    """Same exit codes as the reference command line for the paths it decides before touching CUDA
    (src/main.cpp:99-160): 3 = settings unreadable, 0 = usage."""
    assert _cli(["missing.xml"], tmp_path)[0] == 3
    (tmp_path / "broken.xml").write_text("<Settings><Input><Path inputPath='x'/></Input></Settings>")  # required elements missing
    rc, out = _cli(["broken.xml"], tmp_path)
    assert rc == 3 and "missing" in out
    assert _cli(["a", "b", "c"], tmp_path)[0] == 0                 # wrong argument count: usage text
    rc, out = _cli(["--sequence", 40, 30], tmp_path)
    assert rc == 0 and "Usage" in out
    rc, out = _cli(["--sequence", 40, 30, "out/", "x.raw", "--bogus"], tmp_path)
    assert rc == 0 and "Unknown option" in out
    rc, out = _cli(["--sequence", 40, 30, "out/", "--settings", "missing.xml", "x.raw", "y.raw"], tmp_path)
    assert rc == 3


def test_settings_file_of_the_reference_schema_is_accepted(tmp_path, frames):
    """The reference's settings.xml schema (settings.xml:1-27) with 8-bit frames, through --sequence --settings --list
    (no GPU needed up to the listing)."""
    (frames / "s.xml").write_text(
        '<?xml version="1.0" ?>\n<Settings>\n  <Input>\n    <Path inputPath="./dir8/" />\n'
        '    <Mode Nx="40" Ny="30" imageType="8-bit">\n      <Files file1="f00.raw" file2="f01.raw" />\n    </Mode>\n  </Input>\n'
        '  <Parameters>\n    <Method key="false" />\n    <Solver>\n      <Iterations inner="5" outer="20" />\n'
        '      <Warping levels="20" scaling="0.9" medianRadius="5" />\n'
        '      <Model sigma="0.45" alpha="3.5" e_smooth="0.001" e_data="0.001" />\n    </Solver>\n  </Parameters>\n'
        '  <Output>\n    <Path outputPath="./out/" />\n  </Output>\n</Settings>\n')
    rc, out = _list(["dir8", "--settings", "s.xml"], frames)
    assert rc == 0 and "8 frames of 40x30 (directory, 8-bit)" in out
