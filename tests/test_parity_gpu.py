"""GPU parity: CUDA path (through the C ABI) vs the oracle on seeded synthetic frame pairs.
Tolerance (BASELINE.json north_star): <= 1 LSB per RGB channel, PSNR > 50 dB."""
import pytest

import parity

pytestmark = pytest.mark.gpu


def _need(model):
    if parity.model_dir(model) is None:
        pytest.skip("model %s not shipped to this box" % model)


@pytest.mark.parametrize("w,h", [(256, 256), (640, 360), (96, 64)])
def test_v46_plain(pkg, w, h):
    _need("rife-v4.6")
    res = parity.check_case(pkg, "rife-v4.6", w, h)
    assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, res
    assert res["out_std"] > 5


@pytest.mark.parametrize("t", [0.25, 0.75])
def test_v4_timesteps(pkg, t):
    _need("rife-v4")
    res = parity.check_case(pkg, "rife-v4", 256, 192, timestep=t)
    assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, res


def test_v23_config1(pkg):
    _need("rife-v2.3")
    res = parity.check_case(pkg, "rife-v2.3", 256, 256)
    assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, res


def test_anime_plain(pkg):
    _need("rife-anime")
    res = parity.check_case(pkg, "rife-anime", 256, 192)
    assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, res


@pytest.mark.parametrize("model,tta,ttat", [("rife-v4.6", True, False), ("rife-v4.6", False, True), ("rife-v4.6", True, True),
                                             ("rife-anime", True, True), ("rife-v2.3", True, True), ("rife-v2.3", False, True)])
def test_tta_modes(pkg, model, tta, ttat):
    _need(model)
    res = parity.check_case(pkg, model, 160, 96, tta=tta, tta_temporal=ttat)
    assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, res


@pytest.mark.parametrize("model", ["rife-v2.3", "rife-anime"])
def test_uhd_mode(pkg, model):
    _need(model)
    res = parity.check_case(pkg, model, 256, 192, uhd=True)
    assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, res


def test_timestep_edges_copy_inputs(pkg):
    _need("rife-v4.6")
    import numpy as np
    a, b = parity.synth.pair(64, 64)
    assert np.array_equal(parity.run_gpu(pkg, "rife-v4.6", a, b, 0.0), a)
    assert np.array_equal(parity.run_gpu(pkg, "rife-v4.6", a, b, 1.0), b)


def test_large_motion(pkg):
    _need("rife-v4.6")
    res = parity.check_case(pkg, "rife-v4.6", 640, 352, dx=24, dy=16)
    assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, res
