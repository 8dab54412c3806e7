"""Host emulation of the private layout, the fp16-domain unpack and the mma fragment mapping (tests/emu).
Runs the exact header code the kernels compile (exllamav2_b200/csrc/layout.h, dequant.cuh) with g++ on the CPU."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_layout_and_dequant_emulation(tmp_path):
    exe = tmp_path / "emu_layout"
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(HERE, "emu", "emu_layout.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "layout emulation OK" in r.stdout
