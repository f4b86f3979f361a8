"""GPU-side input pipeline: the reference's transform stack and tuple gather as CUDA kernels.

Mirror of /root/reference/scripts/train.py:119-128 (and scripts/eval.py:97-101):

    tforms = [transforms.Resize(256)]
    if color_jitter > 0: tforms.append(transforms.ColorJitter(brightness=cj, contrast=cj, saturation=cj, hue=0.5))
    tforms += [transforms.ToTensor(), transforms.Normalize(mean=stats[0], std=np.sqrt(stats[1]))]

applied to a whole batch of equally sized uint8 frames that are already in device memory (decode is out of scope):
``ImagePipeline(stats_mean, stats_std)(frames_u8)`` returns the float32 ``[N,3,H',W']`` tensor ``PoseNet.forward``
takes; ``index=`` cuts the batch out of a device-resident sequence (the MF / MFOnline tuple gather,
dataset_loaders/composite.py:60-97,117-126; index arithmetic in geomapnet_b200.data.tuples); ``jitter=`` applies one
ColorJitter draw per image (``ColorJitterSampler`` consumes the torch RNG exactly as torchvision's ``get_params``).
The resize reproduces Pillow's 8-bit bilinear resampling, ColorJitter Pillow's ImageEnhance / HSV arithmetic and
ToTensor / Normalize torchvision's float32 arithmetic BIT FOR BIT (geomapnet_b200/csrc/preprocess_core.h,
jitter_core.h; tests/test_preprocess_host.py runs that arithmetic on the CPU, tests/test_gpu_preprocess.py the
kernels on the GPU, both against Pillow / torchvision).

No CPU path: CPU tensors raise.
"""
import ctypes

import numpy as np
import torch

from .. import _lib

__all__ = ["ImagePipeline", "ColorJitterSampler", "resize_output_size"]


def resize_output_size(H, W, size=256):
    """torchvision.transforms.Resize(size: int): the shorter side becomes `size`."""
    if W <= H:
        return int(size * H / W), size
    return size, int(size * W / H)


class ColorJitterSampler(object):
    """The random part of torchvision.transforms.ColorJitter (scripts/train.py:123-124): per image a random order of
    the four adjustments and a factor for each, drawn from the GLOBAL torch RNG with the same calls in the same order as
    ``ColorJitter.get_params`` -- ``torch.randperm(4)``, then one ``torch.empty(1).uniform_(lo, hi)`` per enabled
    adjustment -- so that a seeded run draws the same augmentations as the reference's transform would.
    Ranges follow ColorJitter._check_input: value v -> [max(0, 1 - v), 1 + v] (hue: [-v, v], v <= 0.5)."""

    def __init__(self, brightness=0.0, contrast=0.0, saturation=0.0, hue=0.0):
        def rng(v, center, bound=None, clip0=True):
            if isinstance(v, (tuple, list)):
                lo, hi = float(v[0]), float(v[1])
            else:
                if v < 0:
                    raise ValueError("ColorJitter value must be non negative")
                lo, hi = center - float(v), center + float(v)
                if clip0:
                    lo = max(lo, 0.0)
            if bound is not None and not (bound[0] <= lo <= hi <= bound[1]):
                raise ValueError("hue values should be between %s" % (bound,))
            return None if lo == hi == center else (lo, hi)
        self.brightness = rng(brightness, 1.0)
        self.contrast = rng(contrast, 1.0)
        self.saturation = rng(saturation, 1.0)
        self.hue = rng(hue, 0.0, bound=(-0.5, 0.5), clip0=False)

    def sample(self, n):
        """-> uint8 tensor [n, 32] (CPU): n records {int32 order[4]; float32 factor[4]} as mapnet_preprocess_run_ex reads
        them.  A disabled adjustment keeps its slot in the order with the neutral factor (1, or 0 for hue)."""
        rec = np.zeros((n, 8), dtype=np.int32)
        fac = rec[:, 4:].view(np.float32)
        for i in range(n):
            order = torch.randperm(4)
            vals = []
            for r, neutral in ((self.brightness, 1.0), (self.contrast, 1.0), (self.saturation, 1.0), (self.hue, 0.0)):
                vals.append(neutral if r is None else float(torch.empty(1).uniform_(r[0], r[1])))
            rec[i, :4] = order.numpy()
            fac[i] = np.asarray(vals, dtype=np.float32)
        return torch.from_numpy(rec.view(np.uint8).reshape(n, 32))


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

    def __call__(self, frames, index=None, jitter=None, return_u8=False):
        """frames: uint8 CUDA tensor [L,H,W,3] (RGB, as PIL decodes).  index: None (every frame, in order) or an int32
        tensor / array of n frame numbers (the tuple gather; geomapnet_b200.data.tuples.batch_frame_indices).
        jitter: None or the [n,32] record tensor of ColorJitterSampler.sample(n).  Returns float32 [n,3,H',W']
        (and the uint8 image that went into ToTensor, [n,H',W',3], when return_u8)."""
        if not frames.is_cuda:
            raise RuntimeError("geomapnet_b200.data.ImagePipeline has no CPU path: frames must be a CUDA tensor")
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
            raise ValueError("expected uint8 frames [N,H,W,3], got %s %s" % (frames.dtype, tuple(frames.shape)))
        frames = frames.contiguous()
        L, H, W, _ = frames.shape
        N = L
        idx_dev = None
        if index is not None:
            idx = torch.as_tensor(np.asarray(index.cpu() if torch.is_tensor(index) else index), dtype=torch.int32)
            if idx.dim() != 1 or idx.numel() < 1:
                raise ValueError("index must be a non-empty 1-D list of frame numbers")
            if int(idx.min()) < 0 or int(idx.max()) >= L:
                raise IndexError("frame index outside [0, %d)" % L)
            N = idx.numel()
            idx_dev = idx.to(frames.device)
        jit_dev = None
        if jitter is not None:
            if tuple(jitter.shape) != (N, 32) or jitter.dtype != torch.uint8:
                raise ValueError("jitter must be the uint8 [n,32] tensor of ColorJitterSampler.sample(n), n=%d" % N)
            jit_dev = jitter.to(frames.device).contiguous()
        Ho, Wo = resize_output_size(H, W, self.size)
        h = self._plan(frames.device, H, W, N)
        out = torch.empty(N, 3, Ho, Wo, dtype=torch.float32, device=frames.device)
        u8 = torch.empty(N, Ho, Wo, 3, dtype=torch.uint8, device=frames.device) if return_u8 else None
        with torch.cuda.device(frames.device):
            _lib.check(_lib.lib().mapnet_preprocess_run_ex(
                h, frames.data_ptr(), N, idx_dev.data_ptr() if idx_dev is not None else None,
                jit_dev.data_ptr() if jit_dev is not None else None,
                self.mean.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                self.std.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), out.data_ptr(),
                u8.data_ptr() if u8 is not None else None, _lib.stream_ptr()), "mapnet_preprocess_run_ex")
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
