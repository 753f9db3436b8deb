"""rife::Combiner (csrc/combiner.h), the flat-combining queue behind option "combine": host-only C++ test with a fake
batch function (tests/emu/test_combiner.cpp), once plain and once under ThreadSanitizer when the toolchain has it."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tsan", [False, True])
def test_combiner(tmp_path, tsan):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "test_combiner")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread", "-I" + os.path.join(ROOT, "rife-ncnn-vulkan_b200", "csrc"),
           os.path.join(ROOT, "tests", "emu", "test_combiner.cpp"), "-o", exe]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 and tsan:
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stdout[-3000:]
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    if tsan and r.returncode != 0 and "ThreadSanitizer" not in r.stdout and "COMBINER" not in r.stdout:
        pytest.skip("ThreadSanitizer cannot run in this sandbox: " + r.stdout[-200:])
    assert r.returncode == 0 and "COMBINER OK" in r.stdout and "WARNING: ThreadSanitizer" not in r.stdout, r.stdout[-3000:]
