"""C++ host program (adanerf_b200/csrc/host): builds with plain g++ against the C ABI; fails loudly without a
GPU; on the B200 box renders an export directory end to end."""
import os
import re
import subprocess

import pytest
import torch

from adanerf_b200 import onnx_weights as ow
from oracle import adanerf_oracle as orc


@pytest.fixture(scope="module")
def viewer():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(g.VIEWER)
    return g.VIEWER


def _export(tmp_path):
    sd0, sd1 = orc.make_weights("shaped", seed=0)
    d = tmp_path / "export"
    ow.write_export_dir(str(d), orc.SCENE_BARBERSHOP, sd0, sd1, 0.2, 8)
    return str(d)


def test_viewer_fails_loudly_without_gpu(viewer, tmp_path):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([viewer, _export(tmp_path), "-f", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1
    assert "K = 8" in r.stdout and "no usable sm_100 device" in r.stderr


def test_viewer_rejects_missing_dir(viewer, tmp_path):
    r = subprocess.run([viewer, str(tmp_path / "nope")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "couldn't read export directory" in r.stderr


@pytest.mark.gpu
def test_viewer_renders_export_dir(viewer, tmp_path):
    d = _export(tmp_path)
    r = subprocess.run([viewer, d, "-s", "400", "300", "-f", "5", "-w"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"5 frames 400x300: ([0-9.]+) ms/frame .*\(([0-9.]+) per ray", r.stdout)
    assert m, r.stdout
    assert 1.0 <= float(m.group(2)) <= 8.0
    ppm = open(os.path.join(d, "adn_frame.ppm"), "rb").read()
    assert ppm.startswith(b"P6\n400 300\n255\n") and len(ppm) == len(b"P6\n400 300\n255\n") + 400 * 300 * 3
