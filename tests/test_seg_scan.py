"""Host logic on CPU: the engine's pass over one segment of the pair list (sequencealigning_b200/csrc/seg_scan.h) --
the threaded scan of a streamed call's first segments must equal the one-thread scan and a plain restatement."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_threaded_segment_scan_equals_the_plain_one(tmp_path):
    exe = str(tmp_path / "seg_scan_check")
    src = os.path.join(ROOT, "tests", "cpp", "seg_scan_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, src], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok "), r.stdout
