// Hardware probe for the strict (split-operand) tensor-core mode: which tcgen05.mma kind::f16 behaviours can the
// conv engines rely on?  One CTA, one 128 x 64 x 64 tile, operands written to shared memory by hand in the
// K-major 128B-swizzle layout the engines use.
//   1. bf16 x bf16 (sanity)               4. fp16 subnormal operands honoured?
//   2. fp16 x fp16                         5. scale-input-d (D = A*B + D * 2^-s)
//   3. A fp16 x B bf16 (mixed formats)     6. accumulation error over 64 k-blocks vs fp64
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I geomapnet_b200/csrc tools/experiments/mma_probe.cu -o gpurun_out/mma_probe
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_ptx.cuh"

using namespace mapnet::ptx;

static __host__ __device__ inline uint32_t idesc(uint32_t afmt, uint32_t bfmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (afmt << 7) | (bfmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void mma_scaled11(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t id) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p, 11;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(id) : "memory");
}

// A: [128][64] 16-bit, B: [64][64] 16-bit (row = n, col = k), both already swizzled by the host.
// mode bit0: second MMA group with scale-input-d = 11 over the same operands;  reps: k-block repetitions
__global__ void __launch_bounds__(128, 1) k_probe(const uint16_t* A, const uint16_t* B, float* D, uint32_t id, int reps, int mode) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 * 64 + 64 * 64) / 8; i += 128)
    reinterpret_cast<uint4*>(sm)[i] = (i < 128 * 64 / 8) ? reinterpret_cast<const uint4*>(A)[i] : reinterpret_cast<const uint4*>(B)[i - 128 * 64 / 8];
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy smem writes -> async proxy (UMMA)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 0 && elect_one()) {
    const uint64_t DB = make_smem_desc_base(16, 1024);
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < 4; ++k)
        mma_bf16(tm, smem_desc(DB, base + k * 32), smem_desc(DB, base + 128 * 128 + k * 32), id, (r > 0 || k > 0) ? 1u : 0u);
    if (mode & 1) {
      mma_scaled11(tm, smem_desc(DB, base), smem_desc(DB, base + 128 * 128), id);
      for (int k = 1; k < 4; ++k)
        mma_bf16(tm, smem_desc(DB, base + k * 32), smem_desc(DB, base + 128 * 128 + k * 32), id, 1u);
    }
    mma_commit(smem_u32(&bar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 64 + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

static size_t swz(int r, int k) { return (size_t)(r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2; }

static uint16_t to16(float v, int fmt) {   // fmt 0 fp16, 1 bf16
  if (fmt == 0) { __half h = __float2half_rn(v); uint16_t u; memcpy(&u, &h, 2); return u; }
  __nv_bfloat16 b = __float2bfloat16_rn(v); uint16_t u; memcpy(&u, &b, 2); return u;
}
static float from16(uint16_t u, int fmt) {
  if (fmt == 0) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
  __nv_bfloat16 b; memcpy(&b, &u, 2); return __bfloat162float(b);
}

struct Case { const char* name; int afmt, bfmt, reps, mode; float ascale, bscale; };

int main() {
  Case cases[] = {
      {"1 bf16 x bf16", 1, 1, 1, 0, 1.f, 1.f},
      {"2 fp16 x fp16", 0, 0, 1, 0, 1.f, 1.f},
      // {"3 fp16 x bf16 (mixed)", 0, 1, ...}: measured on B200 -> "an illegal instruction was encountered": A and B of a
      // kind::f16 MMA must have the same element format
      {"4 fp16 subnormal A (|a| ~ 1e-6)", 0, 0, 1, 0, 1e-6f, 1.f},
      {"4b fp16 subnormal A and B products (1e-6 x 1e-2)", 0, 0, 1, 0, 1e-6f, 1e-2f},
      {"5 scale-input-d 2^-11 (D = AB_k0 + 2^-11 AB + AB_k1..3)", 0, 0, 1, 1, 1.f, 1.f},
      {"6 fp16 accumulate 64 k-blocks", 0, 0, 64, 0, 1.f, 1.f},
      {"6b bf16 accumulate 64 k-blocks", 1, 1, 64, 0, 1.f, 1.f},
  };
  uint16_t *dA, *dB; float* dD;
  cudaMalloc(&dA, 128 * 64 * 2); cudaMalloc(&dB, 64 * 64 * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 + 64 * 128 + 1024);
  srand(7);
  for (const Case& c : cases) {
    std::vector<uint8_t> hA(128 * 128), hB(64 * 128);
    std::vector<float> fA(128 * 64), fB(64 * 64);
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 64; ++k) {
        const float v = c.ascale * ((rand() / (float)RAND_MAX) * 2.f - 1.f);
        const uint16_t u = to16(v, c.afmt);
        fA[r * 64 + k] = from16(u, c.afmt);
        memcpy(&hA[swz(r, k)], &u, 2);
      }
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < 64; ++k) {
        const float v = c.bscale * ((rand() / (float)RAND_MAX) * 2.f - 1.f);
        const uint16_t u = to16(v, c.bfmt);
        fB[n * 64 + k] = from16(u, c.bfmt);
        memcpy(&hB[swz(n, k)], &u, 2);
      }
    cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * 64 * 4);
    k_probe<<<1, 128, 128 * 128 + 64 * 128 + 1024>>>(dA, dB, dD, idesc(c.afmt, c.bfmt, 128, 64), c.reps, c.mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-60s CUDA ERROR %s\n", c.name, cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(128 * 64);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0, sumsq = 0, sumabs_terms = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 64; ++n) {
        double ref = 0, ref0 = 0, terms = 0;
        for (int k = 0; k < 64; ++k) {
          const double p = (double)fA[r * 64 + k] * fB[n * 64 + k];
          ref += p; terms += fabs(p);
          if (k < 16) ref0 += p;
        }
        double want = ref * c.reps;
        if (c.mode & 1) want = ref * ldexp(1.0, -11) + ref0 + (ref - ref0);
        const double err = fabs(hD[r * 64 + n] - want);
        if (err > maxerr) maxerr = err;
        if (fabs(want) > maxref) maxref = fabs(want);
        sumsq += err * err; sumabs_terms += terms * c.reps;
      }
    printf("%-60s max|err| %.3e  max|ref| %.3e  rel %.3e  rms err / mean sum|terms| %.3e\n", c.name, maxerr, maxref,
           maxerr / maxref, sqrt(sumsq / (128 * 64)) / (sumabs_terms / (128 * 64)));
  }
  return 0;
}
