// Fused criterion forward + backward in ONE launch: absolute L1 pose loss and the
// relative-pose (VO) term -- calc_vos_simple (MapNet) or calc_vos in log-quaternion
// space (MapNet++), plus d loss/d pred and d loss/d (sax,saq,srx,srq).
// Replaces ~12 (PoseNet) to ~11 000 (MapNet++ fwd+bwd, N=16,T=5) ATen launches of
// /root/reference/common/criterion.py:42-52,76-109,137-184 and
// /root/reference/common/pose_utils.py:234-260.  Latency-bound:
// one block, one thread per pose, warp-shuffle reductions, no atomics.
#include "kernels.h"
#include "loss_core.h"

namespace mapnet {

__global__ void __launch_bounds__(256)
k_loss(int mode, const float* __restrict__ pred, const float* __restrict__ targ, int N, int Tp, int Tt,
       const float* __restrict__ s4, float* __restrict__ loss, float* __restrict__ dpred,
       float* __restrict__ ds4) {
  pdl_prologue();
  __shared__ float red[8][4];
  float s[4] = {s4[0], s4[1], s4[2], s4[3]};
  const losscore::Cfg c = losscore::make_cfg(mode, N, Tp, Tt, s);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int total = N * Tp;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int n = idx / Tp, i = idx - n * Tp;
    float g[6];
    losscore::pose_contrib(c, pred, targ, n, i, g, acc);
#pragma unroll
    for (int k = 0; k < 6; ++k) dpred[(long long)idx * 6 + k] = g[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = warp_sum(acc[k]);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) red[w][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sums[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww)
#pragma unroll
      for (int k = 0; k < 4; ++k) sums[k] += red[ww][k];
    float L, ds[4];
    losscore::finalize(mode, N, Tp, s, sums, &L, ds);
    loss[0] = L;
#pragma unroll
    for (int k = 0; k < 4; ++k) ds4[k] = ds[k];
  }
}

int launch_loss(int mode, const float* pred, const float* targ, int N, int Tp, int Tt, const float* s4,
                float* loss, float* dpred, float* ds4, cudaStream_t st) {
  MN_CHECK(mode >= 0 && mode <= 3, "loss: bad mode %d", mode);
  MN_CHECK(N > 0 && Tp > 0 && Tt > 0, "loss: empty batch (N=%d Tp=%d Tt=%d)", N, Tp, Tt);
  if (mode == losscore::POSENET) MN_CHECK(Tp == 1 && Tt == 1, "loss: posenet mode wants [N,6] tensors");
  if (mode == losscore::MAPNET) MN_CHECK(Tt == Tp, "loss: mapnet mode wants targ [N,T,6] like pred");
  if (mode == losscore::ONLINE) MN_CHECK(Tp % 2 == 0 && Tt == Tp - 1, "loss: online mode wants pred [N,2T,6], targ [N,2T-1,6]");
  if (mode == losscore::ONLINE_GPS) MN_CHECK(Tp % 2 == 0 && Tt == Tp, "loss: online_gps mode wants pred, targ [N,2T,6]");
  MN_LAUNCH(k_loss, 1, 256, 0, st, mode, pred, targ, N, Tp, Tt, s4, loss, dpred, ds4);
  MN_LAUNCH_CHECK();
  return 0;
}

}  // namespace mapnet
