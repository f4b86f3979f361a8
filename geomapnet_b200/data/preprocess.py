"""GPU-side image pre-processing: the reference's transform stack as two CUDA kernels.

Mirror of /root/reference/scripts/train.py:119-128 (and scripts/eval.py:97-101):

    data_transform = transforms.Compose([transforms.Resize(256), transforms.ToTensor(),
                                         transforms.Normalize(mean=stats[0], std=np.sqrt(stats[1]))])

applied to a whole batch of equally sized uint8 frames that are already in device memory (decode is out of scope):
``ImagePipeline(stats_mean, stats_std)(frames_u8)`` returns the float32 ``[N,3,H',W']`` tensor ``PoseNet.forward``
takes.  The resize reproduces Pillow's 8-bit bilinear resampling bit for bit and ToTensor / Normalize reproduce
torchvision's float32 arithmetic exactly (geomapnet_b200/csrc/preprocess_core.h; tests/test_preprocess_host.py runs
that arithmetic on the CPU against Pillow / torchvision).  ColorJitter (train.py:121-126, training-time augmentation)
is not implemented.  STAGED: the CUDA glue has not run on a GPU yet (tests/test_gpu_preprocess.py is opt-in).

No CPU path: CPU tensors raise.
"""
import ctypes

import numpy as np
import torch

from .. import _lib

__all__ = ["ImagePipeline", "resize_output_size"]


def resize_output_size(H, W, size=256):
    """torchvision.transforms.Resize(size: int): the shorter side becomes `size`."""
    if W <= H:
        return int(size * H / W), size
    return size, int(size * W / H)


class ImagePipeline(object):
    def __init__(self, mean, std, size=256):
        """mean, std: 3 values each, as handed to transforms.Normalize (stats.txt row 0 and sqrt(row 1));
        they are rounded to float32 exactly as torchvision does (torch.as_tensor(..., dtype=float32))."""
        self.mean = np.asarray(mean, dtype=np.float64).astype(np.float32).reshape(3).copy()
        self.std = np.asarray(std, dtype=np.float64).astype(np.float32).reshape(3).copy()
        if not (self.std > 0).all():
            raise ValueError("std must be positive, got %s" % (self.std,))
        self.size = int(size)
        self._plans = {}

    def _plan(self, device, H, W, n):
        key = (device.index, H, W)
        ent = self._plans.get(key)
        if ent is None or ent[1] < n:
            if ent is not None:
                _lib.lib().mapnet_preprocess_destroy(ent[0])
            h = ctypes.c_void_p()
            with torch.cuda.device(device):
                _lib.check(_lib.lib().mapnet_preprocess_create(ctypes.byref(h), H, W, self.size, n), "mapnet_preprocess_create")
            ent = (h, n)
            self._plans[key] = ent
        return ent[0]

    def __call__(self, frames, return_u8=False):
        """frames: uint8 CUDA tensor [N,H,W,3] (RGB, as PIL decodes).  Returns float32 [N,3,H',W']
        (and the resized uint8 frames [N,H',W',3] when return_u8)."""
        if not frames.is_cuda:
            raise RuntimeError("geomapnet_b200.data.ImagePipeline has no CPU path: frames must be a CUDA tensor")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
            raise ValueError("expected uint8 frames [N,H,W,3], got %s %s" % (frames.dtype, tuple(frames.shape)))
        frames = frames.contiguous()
        N, H, W, _ = frames.shape
        Ho, Wo = resize_output_size(H, W, self.size)
        h = self._plan(frames.device, H, W, N)
        out = torch.empty(N, 3, Ho, Wo, dtype=torch.float32, device=frames.device)
        u8 = torch.empty(N, Ho, Wo, 3, dtype=torch.uint8, device=frames.device) if return_u8 else None
        with torch.cuda.device(frames.device):
            _lib.check(_lib.lib().mapnet_preprocess_run(
                h, frames.data_ptr(), N, self.mean.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                self.std.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), out.data_ptr(),
                u8.data_ptr() if u8 is not None else None, _lib.stream_ptr()), "mapnet_preprocess_run")
        return (out, u8) if return_u8 else out

    def close(self):
        for h, _ in self._plans.values():
            _lib.lib().mapnet_preprocess_destroy(h)
        self._plans = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
