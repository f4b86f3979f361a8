// Per-pose core of the fused criterion forward+backward (shared by the CUDA
// kernel in loss.cu and the host-only test shim).  One call handles pose (n,i)
// of pred[N,Tp,6]: it writes d loss / d pred[n,i,:] and returns this pose's
// share of the four L1 sums.  No atomics: a pose that takes part in two
// relative-pose pairs recomputes both pairs' backward.
//
// Restates /root/reference/common/criterion.py:
//   PoseNetCriterion.forward :42-52, MapNetCriterion.forward :76-109,
//   MapNetOnlineCriterion.forward :137-184 (T = s[1] / 2 is py2 integer division).
#pragma once
#include "pose_math.h"

namespace losscore {

enum Mode { POSENET = 0, MAPNET = 1, ONLINE = 2, ONLINE_GPS = 3 };

struct Cfg {
  int mode, N, Tp, Tt;     // pred [N,Tp,6], targ [N,Tt,6]
  float cat, caq, crt, crq;  // exp(-s)/count coefficients of the four L1 means
};

struct Counts { float at, aq, rt, rq; };

PM_HD Counts counts(int mode, int N, int Tp) {
  Counts c;
  if (mode == POSENET) { c.at = c.aq = 3.0f * N * Tp; c.rt = c.rq = 0.0f; }
  else if (mode == MAPNET) { c.at = c.aq = 3.0f * N * Tp; c.rt = c.rq = 3.0f * N * (Tp - 1); }
  else {
    int T = Tp / 2;
    c.at = c.aq = 3.0f * N * T;
    if (mode == ONLINE) { c.rt = c.rq = 3.0f * N * (T - 1); }
    else { c.rt = 2.0f * N * T; c.rq = 0.0f; }
  }
  return c;
}

PM_HD posemath::Pose6 ld_pose(const float* p) {
  posemath::Pose6 r;
  r.t = posemath::v3(p[0], p[1], p[2]);
  r.l = posemath::v3(p[3], p[4], p[5]);
  return r;
}

// acc[0..3] += |.| sums of (abs t, abs q, rel t, rel q) owned by this pose.
PM_HD void pose_contrib(const Cfg& c, const float* pred, const float* targ, int n, int i,
                        float* g /*[6] out*/, float* acc /*[4] in/out*/) {
  using namespace posemath;
  const float* p = pred + ((long long)n * c.Tp + i) * 6;
#pragma unroll
  for (int k = 0; k < 6; ++k) g[k] = 0.0f;
  const int T = (c.mode == ONLINE || c.mode == ONLINE_GPS) ? c.Tp / 2 : c.Tp;
  const bool is_abs = (c.mode == POSENET || c.mode == MAPNET) ? true : (i < T);
  if (is_abs) {
    const float* t = targ + ((long long)n * c.Tt + i) * 6;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      float d = p[k] - t[k];
      float coef = (k < 3) ? c.cat : c.caq;
      acc[k < 3 ? 0 : 1] += fabsf(d);
      g[k] += coef * sgnf(d);
    }
  }
  if (c.mode == MAPNET) {
    // calc_vos_simple (pose_utils.py:234-246): V[i] = p[i+1] - p[i]
    if (i + 1 < c.Tp) {       // pair (i, i+1): this pose is p[i] (owner of the sums)
      const float* pn = p + 6;
      const float* t0 = targ + ((long long)n * c.Tt + i) * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        float d = (pn[k] - p[k]) - (t0[6 + k] - t0[k]);
        float coef = (k < 3) ? c.crt : c.crq;
        acc[k < 3 ? 2 : 3] += fabsf(d);
        g[k] -= coef * sgnf(d);
      }
    }
    if (i > 0) {              // pair (i-1, i): this pose is p[i+1]
      const float* pp = p - 6;
      const float* t0 = targ + ((long long)n * c.Tt + i - 1) * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        float d = (p[k] - pp[k]) - (t0[6 + k] - t0[k]);
        float coef = (k < 3) ? c.crt : c.crq;
        g[k] += coef * sgnf(d);
      }
    }
  } else if (c.mode == ONLINE && !is_abs) {
    const int j = i - T;       // index inside the VO half, poses j=0..T-1, pairs 0..T-2
    // targ[:, T:] holds T-1 relative poses (criterion.py:154)
    if (j + 1 < T) {           // pair j: (p_j, p_{j+1}); this pose is p0 and owns the sums
      Pose6 p0 = ld_pose(p), p1 = ld_pose(p + 6);
      Pose6 o = calc_vo_logq(p0, p1);
      const float* tv = targ + ((long long)n * c.Tt + T + j) * 6;
      float d[6] = {o.t.x - tv[0], o.t.y - tv[1], o.t.z - tv[2], o.l.x - tv[3], o.l.y - tv[4], o.l.z - tv[5]};
      acc[2] += fabsf(d[0]) + fabsf(d[1]) + fabsf(d[2]);
      acc[3] += fabsf(d[3]) + fabsf(d[4]) + fabsf(d[5]);
      Pose6 go;
      go.t = v3(c.crt * sgnf(d[0]), c.crt * sgnf(d[1]), c.crt * sgnf(d[2]));
      go.l = v3(c.crq * sgnf(d[3]), c.crq * sgnf(d[4]), c.crq * sgnf(d[5]));
      Pose6 g0, g1;
      calc_vo_logq_bwd(p0, p1, go, &g0, &g1);
      g[0] += g0.t.x; g[1] += g0.t.y; g[2] += g0.t.z; g[3] += g0.l.x; g[4] += g0.l.y; g[5] += g0.l.z;
    }
    if (j > 0) {               // pair j-1: (p_{j-1}, p_j); this pose is p1
      Pose6 p0 = ld_pose(p - 6), p1 = ld_pose(p);
      Pose6 o = calc_vo_logq(p0, p1);
      const float* tv = targ + ((long long)n * c.Tt + T + j - 1) * 6;
      float d[6] = {o.t.x - tv[0], o.t.y - tv[1], o.t.z - tv[2], o.l.x - tv[3], o.l.y - tv[4], o.l.z - tv[5]};
      Pose6 go;
      go.t = v3(c.crt * sgnf(d[0]), c.crt * sgnf(d[1]), c.crt * sgnf(d[2]));
      go.l = v3(c.crq * sgnf(d[3]), c.crq * sgnf(d[4]), c.crq * sgnf(d[5]));
      Pose6 g0, g1;
      calc_vo_logq_bwd(p0, p1, go, &g0, &g1);
      g[0] += g1.t.x; g[1] += g1.t.y; g[2] += g1.t.z; g[3] += g1.l.x; g[4] += g1.l.y; g[5] += g1.l.z;
    }
  } else if (c.mode == ONLINE_GPS && !is_abs) {
    // gps_mode: absolute xy of the last T predictions vs targ[:, T:, :2] (criterion.py:166,173-176)
    const float* tv = targ + ((long long)n * c.Tt + i) * 6;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float d = p[k] - tv[k];
      acc[2] += fabsf(d);
      g[k] += c.crt * sgnf(d);
    }
  }
}

// loss value and d loss / d (sax,saq,srx,srq) from the four |.| sums.
PM_HD void finalize(int mode, int N, int Tp, const float* s, const float* sums, float* loss, float* ds) {
  Counts cn = counts(mode, N, Tp);
  float ex0 = expf(-s[0]), ex1 = expf(-s[1]);
  float At = sums[0] / cn.at, Aq = sums[1] / cn.aq;
  float L = ex0 * At + s[0] + ex1 * Aq + s[1];
  ds[0] = -ex0 * At + 1.0f;
  ds[1] = -ex1 * Aq + 1.0f;
  ds[2] = 0.0f; ds[3] = 0.0f;
  if (mode != POSENET) {
    float ex2 = expf(-s[2]);
    float Rt = sums[2] / cn.rt;
    L += ex2 * Rt + s[2];
    ds[2] = -ex2 * Rt + 1.0f;
    if (mode != ONLINE_GPS) {
      float ex3 = expf(-s[3]);
      float Rq = sums[3] / cn.rq;
      L += ex3 * Rq + s[3];
      ds[3] = -ex3 * Rq + 1.0f;
    }
  }
  *loss = L;
}

PM_HD Cfg make_cfg(int mode, int N, int Tp, int Tt, const float* s) {
  Cfg c; c.mode = mode; c.N = N; c.Tp = Tp; c.Tt = Tt;
  Counts cn = counts(mode, N, Tp);
  c.cat = expf(-s[0]) / cn.at;
  c.caq = expf(-s[1]) / cn.aq;
  c.crt = (mode != POSENET) ? expf(-s[2]) / cn.rt : 0.0f;
  c.crq = (mode == MAPNET || mode == ONLINE) ? expf(-s[3]) / cn.rq : 0.0f;
  return c;
}

}  // namespace losscore
