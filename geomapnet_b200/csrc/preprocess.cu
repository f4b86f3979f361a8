// GPU-side image pre-processing (STAGED: the arithmetic is checked bit-exactly on the
// CPU against Pillow / torchvision through tests/_hostpre.cpp, the kernels below have not run on a GPU yet --
// tests/test_gpu_preprocess.py is opt-in until they have).
//
// Replaces, for a batch of equally sized uint8 frames already in device memory, the per-image CPU transform stack
// of /root/reference/scripts/train.py:119-128 / scripts/eval.py:97-101:
//   transforms.Resize(256) (Pillow bilinear, 8-bit fixed point, two passes) -> ToTensor -> Normalize(mean, sqrt(var)).
// At ~15 k img/s per GPU the reference's 5 DataLoader workers (mapnet.ini:13) fall short by two orders of magnitude.
//
// HBM-bound byte work: a 640x480 frame is 0.92 MB in, 0.26 MB of intermediate, 1.05 MB (fp32 CHW) out.  Pass 1
// (horizontal) and pass 2 (vertical + ToTensor + Normalize) are separate kernels because Pillow rounds to uint8
// between the passes; consecutive threads walk consecutive output columns, so pass-2 stores are fully coalesced
// and pass-1 / pass-2 loads hit the same few source rows from L1/L2.
#include "kernels.h"
#include "preprocess_core.h"

#include "../../include/mapnet_b200.h"

namespace mapnet {

struct PreprocessPlan {
  int Hin, Win, Hout, Wout, ksh, ksv, max_images;
  int *d_bh, *d_kh, *d_bv, *d_kv;     // bounds / fixed-point weights of the two passes
  uint8_t* d_tmp;                     // [max_images][Hin][Wout][3] after the horizontal pass
};

__global__ void __launch_bounds__(256)
k_pre_resize_h(const uint8_t* __restrict__ img, uint8_t* __restrict__ tmp, const int* __restrict__ bounds,
               const int* __restrict__ kk, int ksize, long long rows /* n*Hin */, int Win, int Wout) {
  pdl_prologue();
  const long long total = rows * Wout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % Wout);
    const long long row = i / Wout;
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const uint8_t* src = img + (row * Win + xmin) * 3;
    const int* k = kk + (long long)xx * ksize;
    uint8_t* dst = tmp + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c] = pre_resample(src + c, 3, n, k);
  }
}

__global__ void __launch_bounds__(256)
k_pre_resize_v_norm(const uint8_t* __restrict__ tmp, float* __restrict__ out, uint8_t* __restrict__ out_u8,
                    const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int nimg, int Hin, int Hout,
                    int Wout, float m0, float m1, float m2, float s0, float s1, float s2) {
  pdl_prologue();
  const long long total = (long long)nimg * Hout * Wout;
  const float mean[3] = {m0, m1, m2}, stdev[3] = {s0, s1, s2};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % Wout);
    const long long r = i / Wout;
    const int yy = (int)(r % Hout);
    const long long n = r / Hout;
    const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    const uint8_t* src = tmp + ((n * Hin + ymin) * Wout + xx) * 3;
    const int* k = kk + (long long)yy * ksize;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint8_t u = pre_resample(src + c, (long long)Wout * 3, cnt, k);
      out[((n * 3 + c) * Hout + yy) * Wout + xx] = pre_normalize(u, mean[c], stdev[c]);
      if (out_u8 != nullptr) out_u8[i * 3 + c] = u;
    }
  }
}

static int upload(const std::vector<int>& v, int** d) {
  MN_CUDA(cudaMalloc((void**)d, v.size() * sizeof(int)));
  MN_CUDA(cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace mapnet

using namespace mapnet;

struct mapnet_preprocess {
  PreprocessPlan p;
};

extern "C" {

int mapnet_preprocess_create(mapnet_preprocess_t** out, int Hin, int Win, int size, int max_images) {
  MN_CHECK(out != nullptr && Hin >= 1 && Win >= 1 && size >= 1 && max_images >= 1, "preprocess_create: bad argument");
  *out = nullptr;
  int ndev = 0;
  MN_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "preprocess_create: no CUDA device (this library has no CPU fallback)");
  mapnet_preprocess* h = new mapnet_preprocess();
  PreprocessPlan& p = h->p;
  memset(&p, 0, sizeof(p));
  p.Hin = Hin; p.Win = Win; p.max_images = max_images;
  pre_resize_output_size(Hin, Win, size, &p.Hout, &p.Wout);
  std::vector<int> bh, kh, bv, kv;
  p.ksh = pre_compute_coeffs(Win, p.Wout, bh, kh);
  p.ksv = pre_compute_coeffs(Hin, p.Hout, bv, kv);
  int r = upload(bh, &p.d_bh);
  if (r == 0) r = upload(kh, &p.d_kh);
  if (r == 0) r = upload(bv, &p.d_bv);
  if (r == 0) r = upload(kv, &p.d_kv);
  if (r == 0 && cudaMalloc((void**)&p.d_tmp, (size_t)max_images * Hin * p.Wout * 3) != cudaSuccess) {
    set_last_error("preprocess_create: cudaMalloc of the intermediate image failed");
    r = 1;
  }
  if (r != 0) { mapnet_preprocess_destroy(h); return r; }
  *out = h;
  return 0;
}

int mapnet_preprocess_output_size(const mapnet_preprocess_t* h, int* Hout, int* Wout) {
  MN_CHECK(h != nullptr && Hout != nullptr && Wout != nullptr, "preprocess_output_size: null argument");
  *Hout = h->p.Hout; *Wout = h->p.Wout;
  return 0;
}

int mapnet_preprocess_run(mapnet_preprocess_t* h, const void* img_nhwc_u8, int n, const float* mean3, const float* std3,
                          float* out_nchw, void* out_u8_or_null, void* stream) {
  MN_CHECK(h != nullptr && img_nhwc_u8 != nullptr && mean3 != nullptr && std3 != nullptr && out_nchw != nullptr,
           "preprocess_run: null argument");
  const PreprocessPlan& p = h->p;
  MN_CHECK(n >= 1 && n <= p.max_images, "preprocess_run: %d images outside [1, max_images=%d]", n, p.max_images);
  for (int c = 0; c < 3; ++c) MN_CHECK(std3[c] > 0.f, "preprocess_run: std[%d] = %f must be positive", c, std3[c]);
  cudaStream_t st = (cudaStream_t)stream;
  const long long t1 = (long long)n * p.Hin * p.Wout, t2 = (long long)n * p.Hout * p.Wout;
  const int g1 = (int)((t1 + 255) / 256 < 148LL * 16 ? (t1 + 255) / 256 : 148LL * 16);
  const int g2 = (int)((t2 + 255) / 256 < 148LL * 16 ? (t2 + 255) / 256 : 148LL * 16);
  MN_LAUNCH(k_pre_resize_h, g1, 256, 0, st, (const uint8_t*)img_nhwc_u8, p.d_tmp, p.d_bh, p.d_kh, p.ksh,
            (long long)n * p.Hin, p.Win, p.Wout);
  MN_LAUNCH_CHECK();
  MN_LAUNCH(k_pre_resize_v_norm, g2, 256, 0, st, p.d_tmp, out_nchw, (uint8_t*)out_u8_or_null, p.d_bv, p.d_kv, p.ksv, n,
            p.Hin, p.Hout, p.Wout, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  MN_LAUNCH_CHECK();
  return 0;
}

int mapnet_preprocess_destroy(mapnet_preprocess_t* h) {
  if (h == nullptr) return 0;
  cudaFree(h->p.d_bh); cudaFree(h->p.d_kh); cudaFree(h->p.d_bv); cudaFree(h->p.d_kv); cudaFree(h->p.d_tmp);
  delete h;
  return 0;
}

}  // extern "C"
