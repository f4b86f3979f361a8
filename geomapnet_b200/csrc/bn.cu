// BatchNorm (training statistics) + ReLU + residual add + stem max-pool, forward
// and backward, on NHWC activations.  HBM-bound: 128-bit vectorised accesses,
// fp32 math, per-thread channel-vector accumulators, deterministic two-stage
// reductions (per-block partials -> fp64 finalize), no atomics.
//
// Replaces the library calls behind torchvision BasicBlock / ResNet.forward
// (cuDNN BN fwd-training/bwd, THCUNN threshold, TH add, SpatialDilatedMaxPooling;
// SURVEY.md section 2c) that /root/reference/models/posenet.py:66 runs.
#include "kernels.h"

namespace mapnet {

static const int kEwThreads = 256;

// ---------------------------------------------------------------------------
// per-channel sums over pixels:  out partials [nblk][NACC][C]
//   NACC=2: (sum y, sum y^2)                        -- forward statistics
//   bwd   : (sum g, sum g*y [, sum g*yd])           -- g = dout * [z > 0]
// ---------------------------------------------------------------------------
template <typename T, int MODE>  // MODE 0: stats(y); 1: bwd(dout,zmask,y); 2: bwd with yd
__global__ void __launch_bounds__(kEwThreads)
k_channel_sums(const T* __restrict__ a, const T* __restrict__ zmask, const T* __restrict__ y,
               const T* __restrict__ yd, long long M, int C, float* __restrict__ partials) {
  constexpr int NACC = (MODE == 0) ? 2 : (MODE == 1 ? 2 : 3);
  const int cv = C >> 3;                      // channel vectors per pixel
  const int rows_par = kEwThreads / cv;       // pixels processed in parallel by the block
  const int tx = threadIdx.x % cv, ty = threadIdx.x / cv;
  float acc[NACC][8];
#pragma unroll
  for (int j = 0; j < NACC; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
  if (ty < rows_par) {
    for (long long r = (long long)blockIdx.x * rows_par + ty; r < M; r += (long long)gridDim.x * rows_par) {
      const long long off = r * C + tx * 8;
      if (MODE == 0) {
        Vec8<T> v; v.load(a + off);
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[0][i] += v.v[i]; acc[1][i] += v.v[i] * v.v[i]; }
      } else {
        Vec8<T> g, yy; g.load(a + off); yy.load(y + off);
        if (zmask != nullptr) {
          Vec8<T> z; z.load(zmask + off);
#pragma unroll
          for (int i = 0; i < 8; ++i) g.v[i] = (z.v[i] > 0.f) ? g.v[i] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[0][i] += g.v[i]; acc[1][i] += g.v[i] * yy.v[i]; }
        if (MODE == 2) {
          Vec8<T> y2; y2.load(yd + off);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[2][i] += g.v[i] * y2.v[i];
        }
      }
    }
  }
  // reduce across ty through shared memory (rows_par <= 32)
  extern __shared__ float sm[];               // [rows_par][NACC][C]
  if (ty < rows_par) {
#pragma unroll
    for (int j = 0; j < NACC; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[((size_t)ty * NACC + j) * C + tx * 8 + i] = acc[j][i];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < NACC * C; idx += kEwThreads) {
    float s = 0.f;
    for (int t = 0; t < rows_par; ++t) s += sm[(size_t)t * NACC * C + idx];
    partials[(size_t)blockIdx.x * NACC * C + idx] = s;
  }
}

static int sums_grid(long long M, int C) {
  const int rows_par = kEwThreads / (C >> 3);
  long long want = (M + (long long)rows_par * 16 - 1) / ((long long)rows_par * 16);   // >=16 rows per thread
  long long cap = 148LL * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <typename T>
int launch_channel_sums(int mode, const T* a, const T* zmask, const T* y, const T* yd, long long M, int C,
                        float* partials, int* nblk_out, cudaStream_t st) {
  MN_CHECK(C % 8 == 0 && C >= 8 && (kEwThreads % (C >> 3)) == 0 && C <= 2048, "channel_sums: unsupported C=%d", C);
  const int grid = sums_grid(M, C);
  const int rows_par = kEwThreads / (C >> 3);
  const int nacc = (mode == 2) ? 3 : 2;
  const size_t smem = (size_t)rows_par * nacc * C * sizeof(float);
  if (mode == 0) k_channel_sums<T, 0><<<grid, kEwThreads, smem, st>>>(a, nullptr, nullptr, nullptr, M, C, partials);
  else if (mode == 1) k_channel_sums<T, 1><<<grid, kEwThreads, smem, st>>>(a, zmask, y, nullptr, M, C, partials);
  else k_channel_sums<T, 2><<<grid, kEwThreads, smem, st>>>(a, zmask, y, yd, M, C, partials);
  MN_LAUNCH_CHECK();
  *nblk_out = grid;
  return 0;
}
template int launch_channel_sums<float>(int, const float*, const float*, const float*, const float*, long long, int, float*, int*, cudaStream_t);
template int launch_channel_sums<bf16>(int, const bf16*, const bf16*, const bf16*, const bf16*, long long, int, float*, int*, cudaStream_t);

// ---------------------------------------------------------------------------
// forward finalize: batch mean / biased var -> scale, shift; running-stat update
// (momentum 0.1, unbiased running var, eps 1e-5: torch.nn.BatchNorm2d defaults)
// ---------------------------------------------------------------------------
__global__ void k_bn_fwd_finalize(const float* __restrict__ partials, int nblk, int C, long long M,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  float* __restrict__ run_mean, float* __restrict__ run_var,
                                  float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                  float* __restrict__ scale, float* __restrict__ shift,
                                  int training, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    double s = 0.0, ss = 0.0;
    for (int b = 0; b < nblk; ++b) {
      s += (double)partials[(size_t)b * 2 * C + c];
      ss += (double)partials[(size_t)b * 2 * C + C + c];
    }
    const double m = s / (double)M;
    double var = ss / (double)M - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    const double unbiased = (M > 1) ? var * (double)M / (double)(M - 1) : var;
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unbiased;
  } else {
    mean = run_mean[c];
    invstd = 1.0f / sqrtf(run_var[c] + eps);
  }
  mean_out[c] = mean;
  invstd_out[c] = invstd;
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
}

int launch_bn_fwd_finalize(const float* partials, int nblk, int C, long long M, const float* gamma,
                           const float* beta, float* run_mean, float* run_var, float* mean_out,
                           float* invstd_out, float* scale, float* shift, int training, cudaStream_t st) {
  k_bn_fwd_finalize<<<cdiv(C, 128), 128, 0, st>>>(partials, nblk, C, M, gamma, beta, run_mean, run_var,
                                                  mean_out, invstd_out, scale, shift, training, 1e-5f, 0.1f);
  MN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// forward apply:  z = relu?( scale*y + shift  [+ zres | + scale2*y2 + shift2] )
// ---------------------------------------------------------------------------
template <typename T, int RES>  // RES 0 none, 1 identity tensor, 2 second BN (downsample branch)
__global__ void __launch_bounds__(kEwThreads)
k_bn_apply(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
           const T* __restrict__ res, const float* __restrict__ scale2, const float* __restrict__ shift2,
           T* __restrict__ z, long long nvec, int C, int relu) {
  const int cv = C >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    Vec8<T> v; v.load(y + i * 8);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = v.v[k] * __ldg(scale + c0 + k) + __ldg(shift + c0 + k);
    if (RES == 1) {
      Vec8<T> r; r.load(res + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += r.v[k];
    } else if (RES == 2) {
      Vec8<T> r; r.load(res + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += r.v[k] * __ldg(scale2 + c0 + k) + __ldg(shift2 + c0 + k);
    }
    Vec8<T> w;
#pragma unroll
    for (int k = 0; k < 8; ++k) w.v[k] = relu ? fmaxf(o[k], 0.f) : o[k];
    w.store(z + i * 8);
  }
}

static int ew_grid(long long n) {
  long long g = (n + kEwThreads - 1) / kEwThreads;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

template <typename T>
int launch_bn_apply(const T* y, const float* scale, const float* shift, int res_mode, const T* res,
                    const float* scale2, const float* shift2, T* z, long long M, int C, int relu,
                    cudaStream_t st) {
  const long long nvec = M * (C >> 3);
  const int grid = ew_grid(nvec);
  if (res_mode == 0) k_bn_apply<T, 0><<<grid, kEwThreads, 0, st>>>(y, scale, shift, nullptr, nullptr, nullptr, z, nvec, C, relu);
  else if (res_mode == 1) k_bn_apply<T, 1><<<grid, kEwThreads, 0, st>>>(y, scale, shift, res, nullptr, nullptr, z, nvec, C, relu);
  else k_bn_apply<T, 2><<<grid, kEwThreads, 0, st>>>(y, scale, shift, res, scale2, shift2, z, nvec, C, relu);
  MN_LAUNCH_CHECK();
  return 0;
}
template int launch_bn_apply<float>(const float*, const float*, const float*, int, const float*, const float*, const float*, float*, long long, int, int, cudaStream_t);
template int launch_bn_apply<bf16>(const bf16*, const float*, const float*, int, const bf16*, const float*, const float*, bf16*, long long, int, int, cudaStream_t);

// ---------------------------------------------------------------------------
// stem: z0 = maxpool3x3s2p1( relu( scale*y0 + shift ) ), argmax position kept
// (first maximum in window scan order wins, as torch's max_pool2d does)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kEwThreads)
k_stem_pool(const T* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
            T* __restrict__ z, uint8_t* __restrict__ amax, int B, int H, int W, int Ho, int Wo, int C) {
  const int cv = C >> 3;
  const long long nvec = (long long)B * Ho * Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    long long p = i / cv;
    const int ow = (int)(p % Wo); p /= Wo;
    const int oh = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float best[8]; int bi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; bi[k] = 0; }
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = __ldg(scale + c0 + k); sh[k] = __ldg(shift + c0 + k); }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * 2 - 1 + kh;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * 2 - 1 + kw;
        if (iw < 0 || iw >= W) continue;
        Vec8<T> v; v.load(y + (((long long)b * H + ih) * W + iw) * C + c0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float a = fmaxf(v.v[k] * sc[k] + sh[k], 0.f);
          if (a > best[k]) { best[k] = a; bi[k] = kh * 3 + kw; }
        }
      }
    }
    Vec8<T> o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = best[k];
    o.store(z + i * 8);
    uint2 packed;
    packed.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
    packed.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
    *reinterpret_cast<uint2*>(amax + i * 8) = packed;
  }
}

template <typename T>
int launch_stem_pool(const T* y, const float* scale, const float* shift, T* z, uint8_t* amax, int B, int H,
                     int W, int Ho, int Wo, int C, cudaStream_t st) {
  const long long nvec = (long long)B * Ho * Wo * (C >> 3);
  k_stem_pool<T><<<ew_grid(nvec), kEwThreads, 0, st>>>(y, scale, shift, z, amax, B, H, W, Ho, Wo, C);
  MN_LAUNCH_CHECK();
  return 0;
}
template int launch_stem_pool<float>(const float*, const float*, const float*, float*, uint8_t*, int, int, int, int, int, int, cudaStream_t);
template int launch_stem_pool<bf16>(const bf16*, const float*, const float*, bf16*, uint8_t*, int, int, int, int, int, int, cudaStream_t);

// stem backward through max-pool and ReLU:  g0[b,ih,iw,c] = [a0>0] * sum_{windows whose argmax is (ih,iw)} dz
template <typename T>
__global__ void __launch_bounds__(kEwThreads)
k_stem_pool_bwd(const T* __restrict__ dz, const uint8_t* __restrict__ amax, const T* __restrict__ y,
                const float* __restrict__ scale, const float* __restrict__ shift, T* __restrict__ g,
                int B, int H, int W, int Ho, int Wo, int C) {
  const int cv = C >> 3;
  const long long nvec = (long long)B * H * W * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    long long p = i / cv;
    const int iw = (int)(p % W); p /= W;
    const int ih = (int)(p % H);
    const int b = (int)(p / H);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    // windows (oh,ow) with 2*oh-1 <= ih <= 2*oh+1
    const int oh_lo = (ih >= 1) ? (ih) / 2 : 0;            // ceil((ih-1)/2)
    const int oh_hi = (ih + 1) / 2;
    const int ow_lo = (iw >= 1) ? (iw) / 2 : 0;
    const int ow_hi = (iw + 1) / 2;
    for (int oh = oh_lo; oh <= oh_hi && oh < Ho; ++oh) {
      const int kh = ih - (oh * 2 - 1);
      if (kh < 0 || kh > 2) continue;
      for (int ow = ow_lo; ow <= ow_hi && ow < Wo; ++ow) {
        const int kw = iw - (ow * 2 - 1);
        if (kw < 0 || kw > 2) continue;
        const long long o = (((long long)b * Ho + oh) * Wo + ow) * C + c0;
        const uint2 pk = *reinterpret_cast<const uint2*>(amax + o);
        Vec8<T> d; d.load(dz + o);
        const int pos = kh * 3 + kw;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t word = (k < 4) ? pk.x : pk.y;
          const int idx = (word >> ((k & 3) * 8)) & 0xff;
          if (idx == pos) acc[k] += d.v[k];
        }
      }
    }
    Vec8<T> yy; yy.load(y + i * 8);
    Vec8<T> o;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float a = yy.v[k] * __ldg(scale + c0 + k) + __ldg(shift + c0 + k);
      o.v[k] = (a > 0.f) ? acc[k] : 0.f;
    }
    o.store(g + i * 8);
  }
}

template <typename T>
int launch_stem_pool_bwd(const T* dz, const uint8_t* amax, const T* y, const float* scale, const float* shift,
                         T* g, int B, int H, int W, int Ho, int Wo, int C, cudaStream_t st) {
  const long long nvec = (long long)B * H * W * (C >> 3);
  k_stem_pool_bwd<T><<<ew_grid(nvec), kEwThreads, 0, st>>>(dz, amax, y, scale, shift, g, B, H, W, Ho, Wo, C);
  MN_LAUNCH_CHECK();
  return 0;
}
template int launch_stem_pool_bwd<float>(const float*, const uint8_t*, const float*, const float*, const float*, float*, int, int, int, int, int, int, cudaStream_t);
template int launch_stem_pool_bwd<bf16>(const bf16*, const uint8_t*, const bf16*, const float*, const float*, bf16*, int, int, int, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------
// backward finalize: d gamma, d beta and the per-channel affine coefficients of
//   dy = A*g + Bc*y + Cc     (g = dout*[z>0];  xhat = (y-mean)*invstd)
// ---------------------------------------------------------------------------
__global__ void k_bn_bwd_finalize(const float* __restrict__ partials, int nblk, int nacc, int which, int C,
                                  long long M, const float* __restrict__ gamma, const float* __restrict__ mean,
                                  const float* __restrict__ invstd, float* __restrict__ dgamma,
                                  float* __restrict__ dbeta, float* __restrict__ coef /*[3][C]*/) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, sy = 0.0;
  for (int b = 0; b < nblk; ++b) {
    s1 += (double)partials[(size_t)b * nacc * C + c];
    sy += (double)partials[(size_t)b * nacc * C + (size_t)which * C + c];
  }
  const double mu = (double)mean[c], is = (double)invstd[c];
  const double s2 = is * (sy - mu * s1);          // sum g * xhat
  dgamma[c] = (float)s2;
  dbeta[c] = (float)s1;
  const double A = (double)gamma[c] * is;
  const double Bc = -A * is * s2 / (double)M;
  const double Cc = -A * s1 / (double)M - Bc * mu;
  coef[c] = (float)A;
  coef[C + c] = (float)Bc;
  coef[2 * C + c] = (float)Cc;
}

int launch_bn_bwd_finalize(const float* partials, int nblk, int nacc, int which, int C, long long M,
                           const float* gamma, const float* mean, const float* invstd, float* dgamma,
                           float* dbeta, float* coef, cudaStream_t st) {
  k_bn_bwd_finalize<<<cdiv(C, 128), 128, 0, st>>>(partials, nblk, nacc, which, C, M, gamma, mean, invstd,
                                                  dgamma, dbeta, coef);
  MN_LAUNCH_CHECK();
  return 0;
}

// backward apply: dy = A*g + B*y + C  [, dyd = Ad*g + Bd*yd + Cd] [, gout = g]
template <typename T, int DS, int GOUT>
__global__ void __launch_bounds__(kEwThreads)
k_bn_bwd_apply(const T* __restrict__ dout, const T* __restrict__ zmask, const T* __restrict__ y,
               const float* __restrict__ coef, T* __restrict__ dy, const T* __restrict__ yd,
               const float* __restrict__ coefd, T* __restrict__ dyd, T* __restrict__ gout, long long nvec,
               int C) {
  const int cv = C >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * 8;
    Vec8<T> g, yy; g.load(dout + i * 8); yy.load(y + i * 8);
    if (zmask != nullptr) {
      Vec8<T> z; z.load(zmask + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) g.v[k] = (z.v[k] > 0.f) ? g.v[k] : 0.f;
    }
    Vec8<T> o;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      o.v[k] = __ldg(coef + c0 + k) * g.v[k] + __ldg(coef + C + c0 + k) * yy.v[k] + __ldg(coef + 2 * C + c0 + k);
    o.store(dy + i * 8);
    if (DS) {
      Vec8<T> y2; y2.load(yd + i * 8);
      Vec8<T> o2;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        o2.v[k] = __ldg(coefd + c0 + k) * g.v[k] + __ldg(coefd + C + c0 + k) * y2.v[k] + __ldg(coefd + 2 * C + c0 + k);
      o2.store(dyd + i * 8);
    }
    if (GOUT) g.store(gout + i * 8);
  }
}

template <typename T>
int launch_bn_bwd_apply(const T* dout, const T* zmask, const T* y, const float* coef, T* dy, const T* yd,
                        const float* coefd, T* dyd, T* gout, long long M, int C, cudaStream_t st) {
  const long long nvec = M * (C >> 3);
  const int grid = ew_grid(nvec);
  const bool ds = (yd != nullptr), go = (gout != nullptr);
  if (ds && !go) k_bn_bwd_apply<T, 1, 0><<<grid, kEwThreads, 0, st>>>(dout, zmask, y, coef, dy, yd, coefd, dyd, gout, nvec, C);
  else if (!ds && go) k_bn_bwd_apply<T, 0, 1><<<grid, kEwThreads, 0, st>>>(dout, zmask, y, coef, dy, yd, coefd, dyd, gout, nvec, C);
  else if (!ds && !go) k_bn_bwd_apply<T, 0, 0><<<grid, kEwThreads, 0, st>>>(dout, zmask, y, coef, dy, yd, coefd, dyd, gout, nvec, C);
  else k_bn_bwd_apply<T, 1, 1><<<grid, kEwThreads, 0, st>>>(dout, zmask, y, coef, dy, yd, coefd, dyd, gout, nvec, C);
  MN_LAUNCH_CHECK();
  return 0;
}
template int launch_bn_bwd_apply<float>(const float*, const float*, const float*, const float*, float*, const float*, const float*, float*, float*, long long, int, cudaStream_t);
template int launch_bn_bwd_apply<bf16>(const bf16*, const bf16*, const bf16*, const float*, bf16*, const bf16*, const float*, bf16*, bf16*, long long, int, cudaStream_t);

}  // namespace mapnet
