"""Image pre-processing arithmetic on the CPU: geomapnet_b200/csrc/preprocess_core.h --
the functions the CUDA kernels call -- compiled for the host (tests/_hostpre.cpp) and compared BIT FOR BIT with the
third-party code the reference's transform stack runs (scripts/train.py:119-128): Pillow's 8-bit bilinear resize and
torchvision's ToTensor / Normalize."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIL = pytest.importorskip("PIL.Image")
T = pytest.importorskip("torchvision.transforms")


@pytest.fixture(scope="module")
def hp():
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhostpre.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "_hostpre.cpp")])
    return ctypes.CDLL(so)


def run_host(hp, img, size, mean32, std32):
    H, W, _ = img.shape
    Ho, Wo = ctypes.c_int(), ctypes.c_int()
    hp.hp_output_size(H, W, size, ctypes.byref(Ho), ctypes.byref(Wo))
    Ho, Wo = Ho.value, Wo.value
    u8 = np.zeros((Ho, Wo, 3), np.uint8)
    f = np.zeros((3, Ho, Wo), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    hp.hp_preprocess(p(img), H, W, Ho, Wo, p(mean32), p(std32), p(u8), p(f))
    return u8, f


# (H, W): 7Scenes 480x640, RobotCar-like 960x1280, portrait, square, tiny (upscaling), odd sizes
SIZES = [(480, 640), (960, 1280), (640, 480), (256, 256), (100, 100), (97, 65), (600, 401), (31, 500), (257, 255)]


@pytest.mark.parametrize("hw", SIZES)
def test_resize_totensor_normalize_bit_exact(hp, hw):
    H, W = hw
    rng = np.random.default_rng(H * 1000 + W)
    img = (rng.random((H, W, 3)) * 256).astype(np.uint8)
    img[: H // 4, : W // 4] = 255; img[-(H // 5):, -(W // 5):] = 0        # saturated regions: the clip8 edges
    stats_mean = rng.random(3) * 0.5 + 0.25                               # float64, as np.loadtxt(stats.txt) gives them
    stats_var = rng.random(3) * 0.08 + 0.01
    tf = T.Compose([T.Resize(256), T.ToTensor(), T.Normalize(mean=stats_mean, std=np.sqrt(stats_var))])
    pil = PIL.fromarray(img)
    ref = tf(pil).numpy()
    u8, f = run_host(hp, img, 256, stats_mean.astype(np.float32), np.sqrt(stats_var).astype(np.float32))
    assert f.shape == ref.shape
    ref_u8 = np.asarray(pil.resize((ref.shape[2], ref.shape[1]), PIL.BILINEAR))
    assert np.array_equal(u8, ref_u8), "resized uint8 image differs from Pillow in %d samples" % int((u8 != ref_u8).sum())
    assert np.array_equal(f.view(np.uint32), ref.view(np.uint32)), "float32 output differs, max abs %g" % float(np.abs(f - ref).max())


def test_other_target_sizes_and_python_size_helper(hp):
    from geomapnet_b200.data.preprocess import resize_output_size
    rng = np.random.default_rng(5)
    for (H, W, size) in [(480, 640, 128), (480, 640, 224), (120, 90, 300), (64, 64, 64)]:
        img = (rng.random((H, W, 3)) * 256).astype(np.uint8)
        mean = np.float32([0.5, 0.5, 0.5]); std = np.float32([0.25, 0.25, 0.25])
        ref = T.Compose([T.Resize(size), T.ToTensor(), T.Normalize(mean=mean.tolist(), std=std.tolist())])(PIL.fromarray(img)).numpy()
        u8, f = run_host(hp, img, size, mean, std)
        assert (f.shape[1], f.shape[2]) == resize_output_size(H, W, size) == (ref.shape[1], ref.shape[2])
        assert np.array_equal(f.view(np.uint32), ref.view(np.uint32))


def test_pipeline_rejects_cpu_tensors_and_bad_input():
    from geomapnet_b200.data import ImagePipeline
    pipe = ImagePipeline([0.5, 0.5, 0.5], [0.2, 0.2, 0.2])
    with pytest.raises(RuntimeError, match="no CPU path"):
        pipe(torch.zeros(2, 48, 64, 3, dtype=torch.uint8))
    with pytest.raises(ValueError):
        ImagePipeline([0.5, 0.5, 0.5], [0.2, 0.0, 0.2])
    # mean / std are rounded to float32 the way torchvision's Normalize does
    p2 = ImagePipeline(np.array([0.1, 0.2, 0.3]), np.sqrt(np.array([0.01, 0.02, 0.03])))
    assert p2.mean.dtype == np.float32 and p2.std[1] == np.float32(np.sqrt(0.02))
