// Image pre-processing arithmetic shared by the CUDA kernels (preprocess.cu) and the CPU parity shim
// (tests/_hostpre.cpp): Resize(256) -> ToTensor -> Normalize, the transform stack of
// /root/reference/scripts/train.py:119-128 and scripts/eval.py:97-101.
//
// The arithmetic that defines the result lives in two third-party packages the reference pins only loosely
// (environment.yml: torchvision, pillow -- unpinned), restated here from their published algorithms:
//  * Pillow `Image.resize(size, BILINEAR)` on 8-bit images (libImaging/Resample.c, unchanged in substance since
//    Pillow 3.x; checked bit-exactly against Pillow 12.2 by tests/test_preprocess_host.py): separable convolution,
//    horizontal pass first, each pass with per-output-pixel windows [xmin, xmin+n) and double-precision triangle
//    weights (support = max(1, in/out)) normalised to sum 1, converted to fixed point with PRECISION_BITS = 22,
//    accumulated in int32 from 1 << 21 and clipped to uint8 BETWEEN the passes;
//  * torchvision `Resize(int)`: the shorter side becomes `size`, the other int(size * long / short);
//    `ToTensor`: uint8 HWC -> float32 CHW / 255;  `Normalize`: (x - mean) / std in float32.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PRE_HD __host__ __device__ __forceinline__
#else
#define PRE_HD inline
#endif

#include <vector>

namespace mapnet {

constexpr int kPrePrecisionBits = 32 - 8 - 2;

PRE_HD uint8_t pre_clip8(int v) {
  v >>= kPrePrecisionBits;          // arithmetic shift, as Pillow's clip8 lookup is indexed
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// one output sample of one pass: window of n source samples `stride` bytes apart, fixed-point weights k
PRE_HD uint8_t pre_resample(const uint8_t* src, long long stride, int n, const int* k) {
  int ss = 1 << (kPrePrecisionBits - 1);
  for (int i = 0; i < n; ++i) ss += (int)src[(long long)i * stride] * k[i];
  return pre_clip8(ss);
}

// ToTensor + Normalize of one sample (mean, stdev already rounded to float32, as torch.as_tensor does)
PRE_HD float pre_normalize(uint8_t u, float mean, float stdev) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), mean), stdev);
#else
  return ((float)u / 255.0f - mean) / stdev;
#endif
}

// torchvision.transforms.Resize(size: int): (H, W) -> (H', W')
inline void pre_resize_output_size(int H, int W, int size, int* Ho, int* Wo) {
  if (W <= H) { *Wo = size; *Ho = (int)((long long)size * H / W); }
  else { *Ho = size; *Wo = (int)((long long)size * W / H); }
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter over the whole axis.
// bounds[2*i] = first source sample of output i, bounds[2*i+1] = window length; kk[i*ksize + j] = weight j.
inline int pre_compute_coeffs(int inSize, int outSize, std::vector<int>& bounds, std::vector<int>& kk) {
  const double scale = (double)inSize / outSize;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;                 // bilinear: support 1
  const int ksize = (int)ceil(support) * 2 + 1;
  bounds.assign((size_t)outSize * 2, 0);
  kk.assign((size_t)outSize * ksize, 0);
  std::vector<double> k((size_t)ksize);
  for (int xx = 0; xx < outSize; ++xx) {
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > inSize) xmax = inSize;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = (a < 1.0) ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (double)(1 << kPrePrecisionBits);
      kk[(size_t)xx * ksize + x] = (k[x] < 0.0) ? (int)(-0.5 + v) : (int)(0.5 + v);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

}  // namespace mapnet
