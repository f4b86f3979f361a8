// Host-only TEST shim: compiles the product's pose_math.h / loss_core.h with g++
// so the hand-derived derivatives are checked on CPU against the reference's
// autograd goldens (tests/test_hostmath.py).  Not part of the product library.
#include "../geomapnet_b200/csrc/loss_core.h"
using namespace posemath;
extern "C" {
void hm_qexp(const float* l, int n, float* out) {
  for (int i = 0; i < n; ++i) { Q4 q = qexp(v3(l[3*i], l[3*i+1], l[3*i+2])); out[4*i]=q.s; out[4*i+1]=q.v.x; out[4*i+2]=q.v.y; out[4*i+3]=q.v.z; }
}
void hm_qlog(const float* q, int n, float* out) {
  for (int i = 0; i < n; ++i) { Q4 a; a.s=q[4*i]; a.v=v3(q[4*i+1],q[4*i+2],q[4*i+3]); V3 o = qlog(a); out[3*i]=o.x; out[3*i+1]=o.y; out[3*i+2]=o.z; }
}
void hm_calc_vos(const float* poses, int N, int T, float* out) {
  for (int n = 0; n < N; ++n) for (int i = 0; i + 1 < T; ++i) {
    const float* p = poses + ((long long)n*T + i)*6;
    Pose6 o = calc_vo_logq(losscore::ld_pose(p), losscore::ld_pose(p+6));
    float* q = out + ((long long)n*(T-1) + i)*6;
    q[0]=o.t.x; q[1]=o.t.y; q[2]=o.t.z; q[3]=o.l.x; q[4]=o.l.y; q[5]=o.l.z;
  }
}
void hm_calc_vos_bwd(const float* poses, int N, int T, const float* w, float* grad) {
  for (long long k = 0; k < (long long)N*T*6; ++k) grad[k] = 0.f;
  for (int n = 0; n < N; ++n) for (int i = 0; i + 1 < T; ++i) {
    const float* p = poses + ((long long)n*T + i)*6;
    const float* g = w + ((long long)n*(T-1) + i)*6;
    Pose6 go; go.t = v3(g[0],g[1],g[2]); go.l = v3(g[3],g[4],g[5]);
    Pose6 g0, g1;
    calc_vo_logq_bwd(losscore::ld_pose(p), losscore::ld_pose(p+6), go, &g0, &g1);
    float* a = grad + ((long long)n*T + i)*6;
    a[0]+=g0.t.x; a[1]+=g0.t.y; a[2]+=g0.t.z; a[3]+=g0.l.x; a[4]+=g0.l.y; a[5]+=g0.l.z;
    a[6]+=g1.t.x; a[7]+=g1.t.y; a[8]+=g1.t.z; a[9]+=g1.l.x; a[10]+=g1.l.y; a[11]+=g1.l.z;
  }
}
void hm_loss(int mode, int N, int Tp, int Tt, const float* pred, const float* targ, const float* s,
             float* loss, float* dpred, float* ds) {
  losscore::Cfg c = losscore::make_cfg(mode, N, Tp, Tt, s);
  float acc[4] = {0,0,0,0};
  for (int n = 0; n < N; ++n) for (int i = 0; i < Tp; ++i)
    losscore::pose_contrib(c, pred, targ, n, i, dpred + ((long long)n*Tp + i)*6, acc);
  losscore::finalize(mode, N, Tp, s, acc, loss, ds);
}
}
