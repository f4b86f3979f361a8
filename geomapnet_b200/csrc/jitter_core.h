// ColorJitter arithmetic on 8-bit RGB pixels, shared by the CUDA kernels (preprocess.cu) and the CPU parity shim
// (tests/_hostpre.cpp).  Training-time augmentation of /root/reference/scripts/train.py:121-126:
//     transforms.ColorJitter(brightness=cj, contrast=cj, saturation=cj, hue=0.5)
// applied between Resize(256) and ToTensor.  torchvision (PIL backend) runs the four adjustments in a random order,
// each one through Pillow; restated here from the published algorithms and checked bit for bit against Pillow 12.2 /
// torchvision 0.26 by tests/test_preprocess_host.py:
//   brightness / contrast / saturation = ImageEnhance.{Brightness,Contrast,Color}(img).enhance(f)
//        = Image.blend(degenerate, img, f) with degenerate = black / solid gray of the ROUNDED MEAN luma / luma image;
//        blend (libImaging/Blend.c): out = (UINT8)(d + f * (p - d)) in float for 0 <= f <= 1, clipped to [0,255] else;
//        luma (libImaging/Convert.c, L24): (19595 R + 38470 G + 7471 B + 0x8000) >> 16
//   hue = RGB -> HSV (8-bit, Convert.c rgb2hsv_row) ; H += uint8(f * 255) mod 256 ; HSV -> RGB (hsv2rgb)
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define JIT_HD __host__ __device__ __forceinline__
#else
#define JIT_HD inline
#endif

namespace mapnet {

enum { JIT_BRIGHTNESS = 0, JIT_CONTRAST = 1, JIT_SATURATION = 2, JIT_HUE = 3 };   // torchvision's fn_idx values

JIT_HD uint8_t jit_luma(uint8_t r, uint8_t g, uint8_t b) {
  return (uint8_t)(((uint32_t)r * 19595u + (uint32_t)g * 38470u + (uint32_t)b * 7471u + 0x8000u) >> 16);
}

// Image.blend(degenerate d, image p, alpha): float32 arithmetic, product and sum rounded separately (no FMA)
JIT_HD uint8_t jit_blend(uint8_t d, uint8_t p, float alpha) {
#if defined(__CUDA_ARCH__)
  const float t = __fadd_rn((float)(int)d, __fmul_rn(alpha, (float)((int)p - (int)d)));
#else
  volatile float prod = alpha * (float)((int)p - (int)d);
  const float t = (float)(int)d + prod;
#endif
  if (alpha >= 0.f && alpha <= 1.0f) return (uint8_t)t;
  if (t <= 0.0f) return 0;
  if (t >= 255.0f) return 255;
  return (uint8_t)t;
}

// Convert.c rgb2hsv_row (the mixed float / double arithmetic of the C source is kept: literals are double)
JIT_HD void jit_rgb2hsv(const uint8_t* in, uint8_t* out) {
  const uint8_t r = in[0], g = in[1], b = in[2];
  const uint8_t maxc = r > g ? (r > b ? r : b) : (g > b ? g : b);
  const uint8_t minc = r < g ? (r < b ? r : b) : (g < b ? g : b);
  uint8_t uh = 0, us = 0;
  if (minc != maxc) {
    const float cr = (float)(maxc - minc);
    const float s = cr / (float)maxc;
    const float rc = ((float)(maxc - r)) / cr;
    const float gc = ((float)(maxc - g)) / cr;
    const float bc = ((float)(maxc - b)) / cr;
    float h;
    if (r == maxc) h = bc - gc;
    else if (g == maxc) h = (float)(2.0 + rc - bc);
    else h = (float)(4.0 + gc - rc);
    h = (float)fmod((h / 6.0 + 1.0), 1.0);
    int ih = (int)(h * 255.0), is = (int)(s * 255.0);
    uh = (uint8_t)(ih < 0 ? 0 : (ih > 255 ? 255 : ih));
    us = (uint8_t)(is < 0 ? 0 : (is > 255 ? 255 : is));
  }
  out[0] = uh; out[1] = us; out[2] = maxc;
}

// Convert.c hsv2rgb
JIT_HD void jit_hsv2rgb(const uint8_t* in, uint8_t* out) {
  const uint8_t h = in[0], s = in[1], v = in[2];
  if (s == 0) { out[0] = v; out[1] = v; out[2] = v; return; }
  const int i = (int)floor((float)h * 6.0 / 255.0);
  const float f = (float)((float)h * 6.0 / 255.0 - (float)i);
  const float fs = (float)(((float)s) / 255.0);
  const int p = (int)round((float)v * (1.0 - fs));
  const int q = (int)round((float)v * (1.0 - fs * f));
  const int t = (int)round((float)v * (1.0 - fs * (1.0 - f)));
  const uint8_t up = (uint8_t)(p < 0 ? 0 : (p > 255 ? 255 : p));
  const uint8_t uq = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
  const uint8_t ut = (uint8_t)(t < 0 ? 0 : (t > 255 ? 255 : t));
  switch (i % 6) {
    case 0: out[0] = v; out[1] = ut; out[2] = up; break;
    case 1: out[0] = uq; out[1] = v; out[2] = up; break;
    case 2: out[0] = up; out[1] = v; out[2] = ut; break;
    case 3: out[0] = up; out[1] = uq; out[2] = v; break;
    case 4: out[0] = ut; out[1] = up; out[2] = v; break;
    default: out[0] = v; out[1] = up; out[2] = uq; break;
  }
}

// one adjustment of one pixel, in place.  gray_mean: the rounded mean luma of the WHOLE image in its state just
// before this adjustment (contrast only).  hue_shift = (uint8_t)(hue_factor * 255), computed by the caller as numpy
// does (np.uint8 of a Python float: truncation toward zero, modulo 256).
JIT_HD void jit_apply(int op, float factor, int gray_mean, uint8_t hue_shift, uint8_t* px) {
  if (op == JIT_BRIGHTNESS) {
    for (int c = 0; c < 3; ++c) px[c] = jit_blend(0, px[c], factor);
  } else if (op == JIT_CONTRAST) {
    for (int c = 0; c < 3; ++c) px[c] = jit_blend((uint8_t)gray_mean, px[c], factor);
  } else if (op == JIT_SATURATION) {
    const uint8_t l = jit_luma(px[0], px[1], px[2]);
    for (int c = 0; c < 3; ++c) px[c] = jit_blend(l, px[c], factor);
  } else {
    uint8_t hsv[3];
    jit_rgb2hsv(px, hsv);
    hsv[0] = (uint8_t)(hsv[0] + hue_shift);
    jit_hsv2rgb(hsv, px);
  }
}

// the per-image parameters of one ColorJitter draw
struct JitterParams {
  int order[4];        // fn_idx: the adjustments in the order they are applied
  float factor[4];     // indexed by adjustment id (brightness, contrast, saturation, hue)
};

JIT_HD uint8_t jit_hue_shift(float hue_factor) {
  // np.uint8(hue_factor * 255): Python float (double) product, C cast to integer (truncation), modulo 256
  const double v = (double)hue_factor * 255.0;
  const long long t = (long long)v;
  return (uint8_t)(t & 0xff);
}

}  // namespace mapnet
