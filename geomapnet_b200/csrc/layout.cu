// Layout transforms around the conv engines:
//  * pack_weights: fp32 master weights in torch layout [Co,Ci,KH,KW] (the
//    state_dict-compatible storage PyTorch owns) -> K-major GEMM
//    operands [Co][KH][KW][Ci] (fprop/wgrad) and [Ci][KH][KW][Co] (dgrad), in the
//    engine's operand type (fp32 or bf16).  All 36 convs in one launch.
//  * unpack_wgrads: wgrad accumulators [Co][KH][KW][Ci] fp32 -> .grad layout.
//  * stem_im2col: NCHW fp32 input -> [B*Ho*Wo][Kpad] patch matrix of the 7x7/s2/p3
//    stem conv (k = (kh*7+kw)*3+ci, zero padded to Kpad), so the stem runs through
//    the same GEMM engines as a 1x1 conv with Ci=Kpad.
#include "kernels.h"

namespace mapnet {

template <typename TW> __device__ __forceinline__ TW cvt_w(float v);
template <> __device__ __forceinline__ float cvt_w<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 cvt_w<bf16>(float v) { return __float2bfloat16_rn(v); }

template <typename TW>
__global__ void k_pack_weights(const WeightDesc* __restrict__ descs, const float* __restrict__ params,
                               TW* __restrict__ w_krsc, TW* __restrict__ w_dg, int round_bf16) {
  const WeightDesc d = descs[blockIdx.y];
  const int KK = d.KH * d.KW;
  const long long total = (long long)d.Co * KK * d.Ci;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(e % d.Ci);
    const int tap = (int)((e / d.Ci) % KK);
    const int co = (int)(e / ((long long)d.Ci * KK));
    float v = 0.f;
    if (d.im2col_k > 0) {
      // packed as [Co][im2col_k]; here KK==1, Ci==im2col_k, ci == k
      const int kreal = d.Ci_real * 49;
      if (ci < kreal) {
        const int c = ci % d.Ci_real, t = ci / d.Ci_real;   // t = kh*7+kw
        v = params[d.p_off + ((long long)co * d.Ci_real + c) * 49 + t];
      }
      if (round_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
      w_krsc[d.k_off + e] = cvt_w<TW>(v);
    } else {
      const int kh = tap / d.KW, kw = tap - kh * d.KW;
      v = params[d.p_off + (((long long)co * d.Ci_real + ci) * d.KH + kh) * d.KW + kw];
      if (round_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
      w_krsc[d.k_off + e] = cvt_w<TW>(v);
      if (w_dg != nullptr) w_dg[d.k_off + ((long long)ci * KK + tap) * d.Co + co] = cvt_w<TW>(v);
    }
  }
}

template <typename TW>
int launch_pack_weights(const WeightDesc* d_descs, int nconv, const float* params, TW* w_krsc, TW* w_dg,
                        int max_elems, int round_bf16, cudaStream_t st) {
  dim3 grid(cdiv(max_elems, 256) < 512 ? cdiv(max_elems, 256) : 512, nconv);
  k_pack_weights<TW><<<grid, 256, 0, st>>>(d_descs, params, w_krsc, w_dg, round_bf16);
  MN_LAUNCH_CHECK();
  return 0;
}
template int launch_pack_weights<float>(const WeightDesc*, int, const float*, float*, float*, int, int, cudaStream_t);
template int launch_pack_weights<bf16>(const WeightDesc*, int, const float*, bf16*, bf16*, int, int, cudaStream_t);

__global__ void k_unpack_wgrads(const WeightDesc* __restrict__ descs, const float* __restrict__ dw,
                                float* __restrict__ grads) {
  const WeightDesc d = descs[blockIdx.y];
  const int KK = d.KH * d.KW;
  if (d.im2col_k > 0) {
    const long long total = (long long)d.Co * d.Ci_real * 49;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
      // e indexes the torch layout [Co][Ci_real][7][7]
      const int t = (int)(e % 49);
      const int c = (int)((e / 49) % d.Ci_real);
      const int co = (int)(e / (49LL * d.Ci_real));
      grads[d.p_off + e] = dw[d.k_off + (long long)co * d.im2col_k + t * d.Ci_real + c];
    }
    return;
  }
  const long long total = (long long)d.Co * KK * d.Ci;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    // e indexes the torch layout [Co][Ci][KH][KW] (coalesced writes)
    const int tap = (int)(e % KK);
    const int ci = (int)((e / KK) % d.Ci);
    const int co = (int)(e / ((long long)KK * d.Ci));
    grads[d.p_off + e] = dw[d.k_off + ((long long)co * KK + tap) * d.Ci + ci];
  }
}

int launch_unpack_wgrads(const WeightDesc* d_descs, int nconv, const float* dw_krsc, float* grads,
                         int max_elems, cudaStream_t st) {
  dim3 grid(cdiv(max_elems, 256) < 512 ? cdiv(max_elems, 256) : 512, nconv);
  k_unpack_wgrads<<<grid, 256, 0, st>>>(d_descs, dw_krsc, grads);
  MN_LAUNCH_CHECK();
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_stem_im2col(const float* __restrict__ x, T* __restrict__ A, int B, int H, int W, int Ho, int Wo, int Kpad) {
  // one thread per (pixel, 8 consecutive k): Kpad % 8 == 0
  const int kv = Kpad >> 3;
  const long long nvec = (long long)B * Ho * Wo * kv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % kv) * 8;
    long long p = i / kv;
    const int ow = (int)(p % Wo); p /= Wo;
    const int oh = (int)(p % Ho);
    const int b = (int)(p / Ho);
    Vec8<T> o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      float v = 0.f;
      if (k < 147) {
        const int c = k % 3, t = k / 3;
        const int kh = t / 7, kw = t - kh * 7;
        const int ih = oh * 2 - 3 + kh, iw = ow * 2 - 3 + kw;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(x + (((long long)b * 3 + c) * H + ih) * W + iw);
      }
      o.v[j] = v;
    }
    o.store(A + i * 8);
  }
}

template <typename T>
int launch_stem_im2col(const float* x_nchw, T* A, int B, int H, int W, int Ho, int Wo, int Kpad, cudaStream_t st) {
  MN_CHECK(Kpad % 8 == 0 && Kpad >= 147, "stem_im2col: bad Kpad");
  const long long nvec = (long long)B * Ho * Wo * (Kpad >> 3);
  long long grid = (nvec + 255) / 256;
  if (grid > 148LL * 32) grid = 148LL * 32;
  k_stem_im2col<T><<<(int)grid, 256, 0, st>>>(x_nchw, A, B, H, W, Ho, Wo, Kpad);
  MN_LAUNCH_CHECK();
  return 0;
}
template int launch_stem_im2col<float>(const float*, float*, int, int, int, int, int, int, cudaStream_t);
template int launch_stem_im2col<bf16>(const float*, bf16*, int, int, int, int, int, int, cudaStream_t);

}  // namespace mapnet
