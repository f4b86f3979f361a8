"""GPU run of the staged image pre-processing kernels (geomapnet_b200/csrc/preprocess.cu) against Pillow /
torchvision, bit for bit.  OPT-IN (MAPNET_STAGED_TESTS=1): the kernels' arithmetic is verified on the CPU
(tests/test_preprocess_host.py) but their launch glue has not run on a GPU yet -- round 1 ended with no GPU
minutes left -- so this test must not gate the round-end suite.  Remove the skip once it has passed on a B200."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MAPNET_STAGED_TESTS") != "1",
                                 reason="staged: set MAPNET_STAGED_TESTS=1 to run the not-yet-GPU-validated kernels")]


@pytest.mark.parametrize("hw,n", [((480, 640), 8), ((97, 65), 3), ((600, 401), 2), ((100, 100), 5)])
def test_gpu_pipeline_matches_pillow_and_torchvision(hw, n):
    from PIL import Image
    import torchvision.transforms as T
    from geomapnet_b200.data import ImagePipeline
    H, W = hw
    rng = np.random.default_rng(H + W + n)
    frames = (rng.random((n, H, W, 3)) * 256).astype(np.uint8)
    stats_mean = rng.random(3) * 0.5 + 0.25
    stats_var = rng.random(3) * 0.08 + 0.01
    tf = T.Compose([T.Resize(256), T.ToTensor(), T.Normalize(mean=stats_mean, std=np.sqrt(stats_var))])
    ref = torch.stack([tf(Image.fromarray(f)) for f in frames]).numpy()
    pipe = ImagePipeline(stats_mean, np.sqrt(stats_var))
    out, u8 = pipe(torch.from_numpy(frames).cuda(), return_u8=True)
    torch.cuda.synchronize()
    ref_u8 = np.stack([np.asarray(Image.fromarray(f).resize((ref.shape[3], ref.shape[2]), Image.BILINEAR)) for f in frames])
    assert np.array_equal(u8.cpu().numpy(), ref_u8)
    got = out.cpu().numpy()
    assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
