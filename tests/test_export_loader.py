"""Export-directory format (config.ini + dataset_info.txt + model{0,1}.onnx): Python and C++ readers, CPU only."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from adanerf_b200 import onnx_weights as ow
from oracle import adanerf_oracle as orc


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from adanerf_b200 import load_library
    return load_library()


def test_onnx_round_trip(tmp_path):
    sd0, sd1 = orc.make_weights("rand", seed=4)
    p = tmp_path / "m.onnx"
    ow.write_onnx_initializers(str(p), {k: v.numpy() for k, v in sd1.items()})
    back = ow.read_onnx_initializers(str(p))
    assert list(back) == list(sd1)
    for k, v in sd1.items():
        np.testing.assert_array_equal(back[k], v.numpy())


def test_cxx_loader_reads_export_dir(lib, tmp_path):
    from adanerf_b200._lib import Scene
    scene = orc.SCENE_PAVILLON
    sd0, sd1 = orc.make_weights("rand", seed=1)
    d = tmp_path / "export"
    ow.write_export_dir(str(d), scene, sd0, sd1, 0.15, 16)
    sc, thr, k, n = Scene(), C.c_float(), C.c_int(), (C.c_int * 2)()
    st = lib.adn_probe_export_dir(str(d).encode(), C.byref(sc), C.byref(thr), C.byref(k), n)
    assert st == 0
    assert abs(thr.value - 0.15) < 1e-7 and k.value == 16
    assert list(n) == [16, 24]                     # initialiser counts of model0 / model1
    np.testing.assert_allclose(list(sc.view_cell_center), scene["view_cell_center"], rtol=1e-6)
    np.testing.assert_allclose(list(sc.depth_range), scene["depth_range"], rtol=1e-6)
    assert abs(sc.fov - scene["fov"]) < 1e-6 and abs(sc.max_depth - scene["max_depth"]) < 1e-5
    assert (sc.n_freq_pos, sc.n_freq_dir) == (10, 4)


def test_cxx_loader_errors(lib, tmp_path):
    assert lib.adn_probe_export_dir(str(tmp_path / "missing").encode(), None, None, None, None) == 5   # ADN_ERR_IO
    d = tmp_path / "bad"
    os.makedirs(d)
    (d / "config.ini").write_text("numRaymarchSamples = [8, 8]\n")
    (d / "dataset_info.txt").write_text("fov = 1.0\n")
    assert lib.adn_probe_export_dir(str(d).encode(), None, None, None, None) == 5


@pytest.mark.skipif(not os.path.isdir("/root/reference/adanerf_real_time_viewer/sample"), reason="reference not mounted")
def test_cxx_loader_reads_shipped_sample(lib):
    from adanerf_b200._lib import Scene
    sc, thr, k, n = Scene(), C.c_float(), C.c_int(), (C.c_int * 2)()
    st = lib.adn_probe_export_dir(b"/root/reference/adanerf_real_time_viewer/sample", C.byref(sc), C.byref(thr), C.byref(k), n)
    assert st == 0 and k.value == 4 and abs(thr.value - 0.15) < 1e-7 and list(n) == [16, 24]
    np.testing.assert_allclose(list(sc.view_cell_center), [2.25, 7.75, 1.5])


@pytest.mark.gpu
def test_render_from_export_dir_matches_state_dict(tmp_path):
    from adanerf_b200 import Renderer
    scene = orc.SCENE_BARBERSHOP
    sd0, sd1 = orc.make_weights("shaped", seed=0)
    d = tmp_path / "export"
    ow.write_export_dir(str(d), scene, sd0, sd1, 0.2, 8)
    r1, thr, K = Renderer.from_export_dir(str(d))
    assert (K, round(thr, 4)) == (8, 0.2)
    r2 = Renderer(scene, sampling_net=sd0, shading_net=sd1)
    pose, rot = torch.tensor(scene["view_cell_center"]), torch.eye(3)
    a = r1.render_camera(pose, rot, 200, 200, thr, K)["rgb"]
    b = r2.render_camera(pose, rot, 200, 200, 0.2, 8)["rgb"]
    assert torch.equal(a, b)
    r1.close()
    r2.close()
