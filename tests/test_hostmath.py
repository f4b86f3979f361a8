"""The product's pose/loss math headers (geomapnet_b200/csrc/pose_math.h,
loss_core.h -- the exact code the CUDA loss kernel runs per thread) compiled for
the host and checked against goldens produced by the REFERENCE's autograd
(common/pose_utils.py, common/criterion.py via oracle.make_goldens)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = {"posenet": 0, "mapnet": 1, "online": 2, "online_gps": 3}


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(HERE, "_build", "libhostmath.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "_hostmath.cpp")])
    return ctypes.CDLL(so)


def fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "pose_math.npz"))


def test_qexp_qlog(hm, gold):
    v = np.ascontiguousarray(gold["qexp_in"], dtype=np.float32)
    q = np.zeros((v.shape[0], 4), np.float32)
    hm.hm_qexp(fp(v), v.shape[0], fp(q))
    np.testing.assert_allclose(q, gold["qexp_out"], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-6)      # property (iv)
    l = np.zeros_like(v)
    qq = np.ascontiguousarray(gold["qexp_out"], dtype=np.float32)
    hm.hm_qlog(fp(qq), v.shape[0], fp(l))
    np.testing.assert_allclose(l, gold["qlog_out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("key", ["n5t3", "n16t5", "n3t2", "n4t7"])
def test_calc_vos_fwd_bwd(hm, gold, key):
    poses = np.ascontiguousarray(gold["vos_in_" + key], dtype=np.float32)
    N, T, _ = poses.shape
    out = np.zeros((N, T - 1, 6), np.float32)
    hm.hm_calc_vos(fp(poses), N, T, fp(out))
    np.testing.assert_allclose(out, gold["vos_" + key], rtol=2e-5, atol=2e-6)
    w = np.ascontiguousarray(gold["vos_w_" + key], dtype=np.float32)
    grad = np.zeros_like(poses)
    hm.hm_calc_vos_bwd(fp(poses), N, T, fp(w), fp(grad))
    ref = gold["vos_grad_" + key]
    assert np.abs(grad - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-6


def test_calc_vos_recovers_delta(hm):
    """known-answer property (v) of SURVEY.md section 4: p_{i+1} = p_i o delta  =>  calc_vos == delta
    (the identity scripts/test_vo.py:33-37 prints)."""
    rng = np.random.default_rng(3)
    from scipy.spatial.transform import Rotation as R
    N, T = 6, 4
    poses = np.zeros((N, T, 6), np.float32)
    deltas = np.zeros((N, T - 1, 6), np.float32)

    def logq(r):
        q = r.as_quat()  # x y z w
        w, v = q[3], q[:3]
        if w < 0:
            w, v = -w, -v
        n = np.linalg.norm(v)
        return v * np.arccos(np.clip(w, -1, 1)) / max(n, 1e-8)

    for n in range(N):
        r = R.from_rotvec(rng.normal(size=3) * 0.3)
        t = rng.normal(size=3)
        poses[n, 0, :3], poses[n, 0, 3:] = t, logq(r)
        for i in range(T - 1):
            dr = R.from_rotvec(rng.normal(size=3) * 0.2)
            dt = rng.normal(size=3) * 0.5
            deltas[n, i, :3], deltas[n, i, 3:] = dt, logq(dr)
            t = t + r.apply(dt)
            r = r * dr
            poses[n, i + 1, :3], poses[n, i + 1, 3:] = t, logq(r)
    out = np.zeros((N, T - 1, 6), np.float32)
    hm.hm_calc_vos(fp(poses), N, T, fp(out))
    np.testing.assert_allclose(out, deltas, rtol=1e-4, atol=2e-5)


def test_degenerate_nan_semantics(hm, gold):
    """identical consecutive rotations: the reference's autograd yields NaN rotation
    gradients (acos'(1) * 0); the product math must be NaN in the same places."""
    poses = np.ascontiguousarray(gold["vos_degen_in"], dtype=np.float32)
    N, T, _ = poses.shape
    out = np.zeros((N, T - 1, 6), np.float32)
    hm.hm_calc_vos(fp(poses), N, T, fp(out))
    np.testing.assert_allclose(out, gold["vos_degen_out"], rtol=1e-5, atol=1e-6)
    grad = np.zeros_like(poses)
    w = np.ones((N, T - 1, 6), np.float32)
    hm.hm_calc_vos_bwd(fp(poses), N, T, fp(w), fp(grad))
    ref = gold["vos_degen_grad"]
    assert np.array_equal(np.isnan(grad), np.isnan(ref))
    m = ~np.isnan(ref)
    np.testing.assert_allclose(grad[m], ref[m], rtol=1e-4, atol=1e-6)


CRIT_KEYS = ["posenet_n64t1", "posenet_n7t1", "mapnet_n32t3", "mapnet_n5t2", "online_n16t10",
             "online_n3t4", "online_gps_n16t10", "online_gps_n2t6"]


@pytest.mark.parametrize("key", CRIT_KEYS)
def test_criteria(hm, gold, key):
    kind = key.rsplit("_n", 1)[0]
    pred = np.ascontiguousarray(gold["crit_pred_" + key], dtype=np.float32)
    targ = np.ascontiguousarray(gold["crit_targ_" + key], dtype=np.float32)
    if kind == "posenet":
        N, Tp, Tt = pred.shape[0], 1, 1
    else:
        N, Tp, Tt = pred.shape[0], pred.shape[1], targ.shape[1]
    s = np.array([0.0, -3.0, 0.0, -3.0], np.float32)
    loss = np.zeros(1, np.float32)
    dpred = np.zeros_like(pred)
    ds = np.zeros(4, np.float32)
    hm.hm_loss(MODES[kind], N, Tp, Tt, fp(pred), fp(targ), fp(s), fp(loss), fp(dpred), fp(ds))
    ref_loss = float(gold["crit_loss_" + key].reshape(-1)[0])
    assert abs(loss[0] - ref_loss) <= 2e-6 * abs(ref_loss) + 1e-6
    ref_d = gold["crit_dpred_" + key]
    assert np.abs(dpred - ref_d).max() <= 1e-4 * np.abs(ref_d).max() + 1e-7
    ref_ds = gold["crit_ds_" + key]
    for i in range(len(ref_ds)):
        if not np.isnan(ref_ds[i]):
            assert abs(ds[i] - ref_ds[i]) <= 2e-6 * abs(ref_ds[i]) + 1e-6


def test_mapnet_reduces_to_posenet(hm):
    """property (vii): with T=1 the VO term is empty (0/0 -> NaN in the reference as
    well); with constant-velocity identical pred/targ VOs it is exactly srx+srq."""
    rng = np.random.default_rng(0)
    N, T = 5, 3
    targ = rng.normal(size=(N, T, 6)).astype(np.float32)
    off = rng.normal(size=(N, 1, 6)).astype(np.float32)
    pred = (targ + off).astype(np.float32)        # same VOs, shifted absolute poses
    s = np.array([0.1, -2.0, 0.3, -1.0], np.float32)
    loss = np.zeros(1, np.float32); dpred = np.zeros_like(pred); ds = np.zeros(4, np.float32)
    hm.hm_loss(1, N, T, T, fp(pred), fp(targ), fp(s), fp(loss), fp(dpred), fp(ds))
    loss_p = np.zeros(1, np.float32); dp2 = np.zeros_like(pred); ds2 = np.zeros(4, np.float32)
    p2 = np.ascontiguousarray(pred.reshape(-1, 6)); t2 = np.ascontiguousarray(targ.reshape(-1, 6))
    hm.hm_loss(0, N * T, 1, 1, fp(p2), fp(t2), fp(s), fp(loss_p), fp(dp2), fp(ds2))
    # VO differences are ~1e-7 rounding noise, so the term is ~ srx + srq
    assert abs(loss[0] - (loss_p[0] + s[2] + s[3])) < 1e-4
