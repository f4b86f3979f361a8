// CPU parity shim for the image pre-processing arithmetic (TEST INFRASTRUCTURE): runs
// geomapnet_b200/csrc/preprocess_core.h -- the same functions the CUDA kernels call -- on host memory so that
// tests/test_preprocess_host.py can compare them bit for bit with Pillow / torchvision.
#include "../geomapnet_b200/csrc/preprocess_core.h"

using namespace mapnet;

extern "C" {

void hp_output_size(int H, int W, int size, int* Ho, int* Wo) { pre_resize_output_size(H, W, size, Ho, Wo); }

// img [H][W][3] uint8 -> out_u8 [Ho][Wo][3] (may be null) and out_f [3][Ho][Wo] (may be null)
void hp_preprocess(const uint8_t* img, int H, int W, int Ho, int Wo, const float* mean3, const float* std3,
                   uint8_t* out_u8, float* out_f) {
  std::vector<int> bh, kh, bv, kv;
  const int ksh = pre_compute_coeffs(W, Wo, bh, kh);
  const int ksv = pre_compute_coeffs(H, Ho, bv, kv);
  std::vector<uint8_t> tmp((size_t)H * Wo * 3);
  for (int y = 0; y < H; ++y)
    for (int xx = 0; xx < Wo; ++xx)
      for (int c = 0; c < 3; ++c)
        tmp[((size_t)y * Wo + xx) * 3 + c] =
            pre_resample(img + ((size_t)y * W + bh[2 * xx]) * 3 + c, 3, bh[2 * xx + 1], &kh[(size_t)xx * ksh]);
  for (int yy = 0; yy < Ho; ++yy)
    for (int xx = 0; xx < Wo; ++xx)
      for (int c = 0; c < 3; ++c) {
        const uint8_t u = pre_resample(&tmp[((size_t)bv[2 * yy] * Wo + xx) * 3 + c], (long long)Wo * 3, bv[2 * yy + 1],
                                       &kv[(size_t)yy * ksv]);
        if (out_u8) out_u8[((size_t)yy * Wo + xx) * 3 + c] = u;
        if (out_f) out_f[((size_t)c * Ho + yy) * Wo + xx] = pre_normalize(u, mean3[c], std3[c]);
      }
}

}  // extern "C"

// ---- ColorJitter arithmetic (geomapnet_b200/csrc/jitter_core.h) ------------------------------------------------
#include "../geomapnet_b200/csrc/jitter_core.h"

extern "C" {

// rounded mean luma of an RGB image: int(ImageStat.Stat(img.convert("L")).mean[0] + 0.5)
int hj_gray_mean(const uint8_t* img, long long npix) {
  unsigned long long s = 0;
  for (long long i = 0; i < npix; ++i) s += jit_luma(img[3 * i], img[3 * i + 1], img[3 * i + 2]);
  return (int)((double)s / (double)npix + 0.5);
}

// one adjustment of a whole image, in place
void hj_apply(uint8_t* img, long long npix, int op, float factor) {
  const int mean = (op == JIT_CONTRAST) ? hj_gray_mean(img, npix) : 0;
  const uint8_t shift = jit_hue_shift(factor);
  for (long long i = 0; i < npix; ++i) jit_apply(op, factor, mean, shift, img + 3 * i);
}

void hj_rgb2hsv(const uint8_t* in, uint8_t* out, long long npix) { for (long long i = 0; i < npix; ++i) jit_rgb2hsv(in + 3 * i, out + 3 * i); }
void hj_hsv2rgb(const uint8_t* in, uint8_t* out, long long npix) { for (long long i = 0; i < npix; ++i) jit_hsv2rgb(in + 3 * i, out + 3 * i); }

}  // extern "C"
