"""CPU restatement of the reference's pose-graph optimisation -- TEST INFRASTRUCTURE.

Follows /root/reference/common/pose_utils.py:458-804 (PoseGraph, PoseGraphFC, optimize_poses) and the numpy helpers
it calls (:370-456: skew, dpq_q, dpsq_q, dpsq_p, dqstq_t, m_rot) line by line, in numpy float64.  The quaternion
helpers come from a third-party package the reference pins only by name (`transforms3d`, environment.yml:19; imported as
txq at pose_utils.py:13-14) that is not installed here: `qmult`, `qconjugate`, `qinverse`, `rotate_vector` are restated
from its published source (transforms3d/quaternions.py).  Pinned against the reference's own classes executed from
/root/reference (``load_reference()`` below; tests/test_pgo_oracle.py, build container) and against committed goldens
(tests/golden/pgo.npz, made by oracle/make_goldens.py from the reference).  Only tests/ may import this.
"""
import math
import os
import types

import numpy as np

REF = "/root/reference/common/pose_utils.py"


# ---- transforms3d.quaternions (published algorithms) ----------------------------------------------------------------
def qmult(q1, q2):
    w1, x1, y1, z1 = q1
    w2, x2, y2, z2 = q2
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2])


def qconjugate(q):
    return np.array(q) * np.array([1.0, -1, -1, -1])


def qinverse(q):
    return qconjugate(q) / np.dot(q, q)


def rotate_vector(v, q):
    varr = np.zeros((4,))
    varr[1:] = v
    return qmult(q, qmult(varr, qconjugate(q)))[1:]


# ---- pose_utils.py:370-456 ---------------------------------------------------------------------------------------------
def skew(x):
    x = np.asarray(x).reshape(3)
    return np.asarray([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]])


def dpq_q(p):
    p = np.asarray(p).reshape(4)
    J = np.zeros((4, 4))
    J[0, 0] = p[0]; J[0, 1:] = -p[1:]; J[1:, 0] = p[1:]; J[1:, 1:] = p[0] * np.eye(3) + skew(p[1:])
    return J


def dpsq_q(p):
    p = np.asarray(p).reshape(4)
    J = np.zeros((4, 4))
    J[0, 0] = p[0]; J[0, 1:] = -p[1:]; J[1:, 0] = -p[1:]; J[1:, 1:] = p[0] * np.eye(3) - skew(p[1:])
    return J


def dpsq_p(q):
    q = np.asarray(q).reshape(4)
    J = np.zeros((4, 4))
    J[0, 0] = q[0]; J[0, 1:] = q[1:]; J[1:, 0] = q[1:]; J[1:, 1:] = -q[0] * np.eye(3) + skew(q[1:])
    return J


def dqstq_t(q):
    q = np.asarray(q).reshape(4)
    v = q[1:].reshape(3, 1)
    return (q[0] * q[0] - float((v.T @ v)[0, 0])) * np.eye(3) + 2 * (v @ v.T) - 2 * q[0] * skew(q[1:])


def m_rot(x):
    return dpq_q(x) @ np.vstack((np.zeros((1, 3)), np.eye(3)))


# ---- pose_utils.py:458-773 ---------------------------------------------------------------------------------------------
def edges(N, fc):
    return [(i, j) for i in range(N) for j in range(i + 1, N)] if fc else [(i, i + 1) for i in range(N - 1)]


def optimize(poses, vos, sax=1, saq=1, srx=1, srq=1, n_iters=10, fc=False, exact_solve=False):
    """PoseGraph.optimize (:577-613) / PoseGraphFC.optimize (:737-773): poses [N,7], vos [E,7] -> [N,7].

    The linear solve follows the reference LITERALLY (:605-608):
        R = slin.cholesky(H)                     # upper, H = R' R
        y = slin.solve_triangular(R.T, -b)       # <- default lower=False: LAPACK reads only the UPPER triangle of R.T,
        x = slin.solve_triangular(R, y)          #    i.e. its diagonal, so y = -b / diag(R)
    so the step is x = R^-1 diag(R)^-1 (-b), not the Gauss-Newton step H^-1 (-b).  That is what scripts/eval.py computes
    and what the published numbers (README.md:112-180) were made with; exact_solve=True does the forward substitution."""
    poses = np.asarray(poses, dtype=np.float64)
    vos = np.asarray(vos, dtype=np.float64)
    N = len(poses)
    z = poses.copy().reshape(-1)
    L_ax, L_aq = np.eye(3) / math.sqrt(sax), np.eye(4) / math.sqrt(saq)      # cholesky(I / s).T
    L_rx, L_rq = np.eye(3) / math.sqrt(srx), np.eye(4) / math.sqrt(srq)
    E = edges(N, fc)
    for _ in range(n_iters):
        rows, res = [], []
        for i in range(N):                                                    # unary (:473-483, :516-523)
            jt = np.zeros((3, 6 * N)); jt[:, 6 * i:6 * i + 3] = np.eye(3)
            rows.append(L_ax @ jt); res.append(L_ax @ (z[7 * i:7 * i + 3] - poses[i, :3]))
            jr = np.zeros((4, 6 * N)); jr[:, 6 * i + 3:6 * i + 6] = m_rot(z[7 * i + 3:7 * i + 7])
            rows.append(L_aq @ jr); res.append(L_aq @ (z[7 * i + 3:7 * i + 7] - poses[i, 3:]))
        for k, (i, j) in enumerate(E):                                        # pairwise (:486-505, :526-549)
            qi, qj = z[7 * i + 3:7 * i + 7], z[7 * j + 3:7 * j + 7]
            dt = dqstq_t(qi)
            jt = np.zeros((3, 6 * N)); jt[:, 6 * i:6 * i + 3] = -dt; jt[:, 6 * j:6 * j + 3] = dt
            rt = rotate_vector(z[7 * j:7 * j + 3] - z[7 * i:7 * i + 3], qinverse(qi)) - vos[k, :3]
            rows.append(L_rx @ jt); res.append(L_rx @ rt)
            jr = np.zeros((4, 6 * N))
            jr[:, 6 * i + 3:6 * i + 6] = dpsq_p(qj) @ m_rot(qi)
            jr[:, 6 * j + 3:6 * j + 6] = dpsq_q(qi) @ m_rot(qj)
            rq = qmult(qinverse(qi), qj) - vos[k, 3:]
            rows.append(L_rq @ jr); res.append(L_rq @ rq)
        J, r = np.vstack(rows), np.concatenate(res)
        H, b = J.T @ J, J.T @ r
        R = np.linalg.cholesky(H).T                                           # H = R' R (scipy.linalg.cholesky, upper)
        y = np.linalg.solve(R.T, -b) if exact_solve else -b / np.diag(R)
        x = np.linalg.solve(R, y)
        for i in range(N):                                                    # update_on_manifold (:552-575)
            z[7 * i:7 * i + 3] += x[6 * i:6 * i + 3]
            qm = x[6 * i + 3:6 * i + 6]
            n = np.linalg.norm(qm)
            dq = np.concatenate(([math.cos(n)], np.sinc(n / np.pi) * qm))
            z[7 * i + 3:7 * i + 7] = qmult(z[7 * i + 3:7 * i + 7], dq)
    return z.reshape(-1, 7)


def optimize_poses(pred_poses, vos=None, fc_vos=False, target_poses=None, sax=1, saq=1, srx=1, srq=1):
    """pose_utils.py:775-804"""
    if vos is None:
        if target_poses is None:
            return None
        target_poses = np.asarray(target_poses, dtype=np.float64)
        vos = np.zeros((len(target_poses) - 1, 7))
        for i in range(len(vos)):
            vos[i, :3] = target_poses[i + 1, :3] - target_poses[i, :3]
            vos[i, 3:] = qmult(qinverse(target_poses[i, 3:]), target_poses[i + 1, 3:])
    return optimize(pred_poses, vos, sax, saq, srx, srq, fc=fc_vos)


# ---- the reference's own code, executed from its source (build container only) -----------------------------------------
def available():
    return os.path.exists(REF)


def load_reference():
    """optimize_poses / PoseGraph / PoseGraphFC of /root/reference/common/pose_utils.py, executed from lines 306-804 (the
    numpy section up to the first Python-2 print statement) with `xrange`, and `txq` bound to the restatement above."""
    import scipy.linalg as slin
    src = open(REF).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("## NUMPY"))
    end = next(i for i, l in enumerate(src) if l.startswith("def align_3d_pts"))
    body = "\n".join(src[start:end]).replace("print 'Specify either VO or target poses'", "pass")
    txq = types.SimpleNamespace(qmult=qmult, qinverse=qinverse, rotate_vector=rotate_vector, qconjugate=qconjugate)
    ns = {"np": np, "math": math, "slin": slin, "txq": txq, "xrange": range}
    exec(compile(body, REF, "exec"), ns)
    # NumPy >= 2 no longer converts the size-1 arrays x[k] of a (3,1) column to scalars inside np.asarray([[0, -x[2], ...
    # (pose_utils.py:370-378 was written for NumPy 1.14): same function, column squeezed first
    ref_skew = ns["skew"]
    ns["skew"] = lambda x: ref_skew(np.asarray(x).reshape(3))
    return ns
