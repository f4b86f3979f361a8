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
