"""Writes tests/golden/ba_leaf_ref.npz: seeded inputs and the outputs of the REFERENCE's own leaf arithmetic for the BA path
(oracle/_ref/libref_leaf.so = the unmodified g2o / AirDOS type sources compiled by `make -C oracle ref`, see oracle/ref_leaf.cpp).
Run in the build container (needs /root/reference); the fixture travels to the GPU box, the reference does not.

    python oracle/gen_ref_leaf_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_leaf.so")


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def rand_quat(rng, n, small=False):
    if small:
        v = rng.normal(0, 0.05, (n, 3)); q = np.concatenate([v, np.ones((n, 1))], 1)
    else:
        q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 3] < 0] *= -1          # SE3Quat keeps w >= 0
    return q


def main(n=256, seed=20261017):
    L = C.CDLL(LIB)
    L.ref_rigid_error.restype = C.c_double
    L.ref_rigid_error.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    L.ref_dist_oplus.restype = C.c_double
    L.ref_dist_oplus.argtypes = [C.c_double, C.c_double]
    L.ref_motion_edge.argtypes = [C.c_void_p] * 4 + [C.c_double] + [C.c_void_p] * 4
    rng = np.random.default_rng(seed)
    cam = np.array([772.548, 772.548, 320.0, 240.0, 193.137])
    out = dict(cam=cam)
    # ---- poses, points in front of the camera, observations near the projection
    q = rand_quat(rng, n); t = rng.normal(0, 2.0, (n, 3))
    Xc = np.stack([rng.uniform(-6, 6, n), rng.uniform(-4, 4, n), rng.uniform(1.5, 40, n)], 1)
    Xc[:8, 2] = rng.uniform(-5, -0.5, 8)                        # a few behind the camera (isDepthPositive = false)
    R = np.zeros((n, 3, 3))
    for i in range(n):
        x, y, z, w = q[i]
        R[i] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    X = np.einsum("nji,nj->ni", R, Xc - t)                       # Xw = R^T (Xc - t)
    u = Xc[:, 0] / Xc[:, 2] * cam[0] + cam[2]; v = Xc[:, 1] / Xc[:, 2] * cam[1] + cam[3]
    obs = np.stack([u + rng.normal(0, 2, n), v + rng.normal(0, 2, n), u - cam[4] / Xc[:, 2] + rng.normal(0, 2, n)], 1)
    obs = obs.astype(np.float32).astype(np.float64)             # the map stores float key-points
    obs[:, 2] = np.abs(obs[:, 2])
    Xf = X.astype(np.float32).astype(np.float64)               # MapPoint::GetWorldPos is a float cv::Mat: the OnlyPose edges get float positions
    out.update(pose_q=q, pose_t=t, X=X, Xf=Xf, obs=obs)
    se, sJi, sJj, sdp = np.zeros((n, 3)), np.zeros((n, 9)), np.zeros((n, 18)), np.zeros(n, np.int32)
    me, mJi, mJj, mdp = np.zeros((n, 2)), np.zeros((n, 6)), np.zeros((n, 12)), np.zeros(n, np.int32)
    pe, pJ, pme, pmJ = np.zeros((n, 3)), np.zeros((n, 18)), np.zeros((n, 2)), np.zeros((n, 12))
    for i in range(n):
        dp = C.c_int()
        L.ref_edge_stereo(P(q[i]), P(t[i]), P(X[i]), P(obs[i]), P(cam), P(se[i]), P(sJi[i]), P(sJj[i]), C.byref(dp)); sdp[i] = dp.value
        L.ref_edge_mono(P(q[i]), P(t[i]), P(X[i]), P(obs[i]), P(cam), P(me[i]), P(mJi[i]), P(mJj[i]), C.byref(dp)); mdp[i] = dp.value
        L.ref_edge_stereo_onlypose(P(q[i]), P(t[i]), P(Xf[i]), P(obs[i]), P(cam), P(pe[i]), P(pJ[i]))
        L.ref_edge_mono_onlypose(P(q[i]), P(t[i]), P(Xf[i]), P(obs[i]), P(cam), P(pme[i]), P(pmJ[i]))
    out.update(stereo_err=se, stereo_Ji=sJi, stereo_Jj=sJj, stereo_depth_positive=sdp, mono_err=me, mono_Ji=mJi, mono_Jj=mJj, mono_depth_positive=mdp,
               onlypose_stereo_err=pe, onlypose_stereo_J=pJ, onlypose_mono_err=pme, onlypose_mono_J=pmJ)
    # ---- VertexSE3Expmap::oplusImpl: generic updates, tiny updates (theta < 1e-5 branch), large rotations
    upd = np.concatenate([rng.normal(0, 0.05, (n, 3)), rng.normal(0, 0.3, (n, 3))], 1)
    upd[:16, :3] = rng.normal(0, 2e-6, (16, 3))                 # Taylor branch of SE3Quat::exp
    upd[16:32, :3] = rng.normal(0, 1.5, (16, 3))                # large angles (Quaterniond(R) non-trace branches)
    oq, ot = np.zeros((n, 4)), np.zeros((n, 3))
    for i in range(n):
        L.ref_pose_oplus(P(q[i]), P(t[i]), P(upd[i]), P(oq[i]), P(ot[i]))
    out.update(pose_update=upd, pose_oplus_q=oq, pose_oplus_t=ot)
    # ---- Converter::toSE3Quat on float rotation matrices (all four branches of Quaterniond(R)), to_homogeneous_matrix
    Rf = R.astype(np.float32).astype(np.float64); tf = t.astype(np.float32).astype(np.float64)
    cq, ct, T16 = np.zeros((n, 4)), np.zeros((n, 3)), np.zeros((n, 16))
    for i in range(n):
        L.ref_pose_from_rt(P(np.ascontiguousarray(Rf[i])), P(tf[i]), P(cq[i]), P(ct[i]))
        L.ref_pose_to_matrix(P(q[i]), P(t[i]), P(T16[i]))
    out.update(conv_R=Rf, conv_t=tf, conv_q=cq, conv_tt=ct, pose_matrix=T16)
    # ---- rigidity edge, bone length / point vertices
    p1 = rng.normal(0, 1.0, (n, 3)) + [0, 0, 8]; p2 = p1 + rng.normal(0, 0.3, (n, 3)); d = np.linalg.norm(p1 - p2, axis=1) + rng.normal(0, 0.02, n)
    rig = np.array([L.ref_rigid_error(P(p1[i]), P(p2[i]), d[i]) for i in range(n)])
    dup = rng.normal(0, 0.01, n)
    dist_o = np.array([L.ref_dist_oplus(d[i], dup[i]) for i in range(n)])
    d3 = rng.normal(0, 0.1, (n, 3)); xo = np.zeros((n, 3))
    for i in range(n):
        L.ref_point_oplus(P(X[i]), P(d3[i]), P(xo[i]))
    out.update(joint_a=p1, joint_b=p2, bone=d, rigid_err=rig, bone_update=dup, bone_oplus=dist_o, point_update=d3, point_oplus=xo)
    # ---- motion edge and VertexSE3::oplusImpl
    mq = rand_quat(rng, n, small=True); mt = rng.normal(0, 0.5, (n, 3)); dt = rng.choice([1.0, 0.5, 2.0], n)
    mer, J1, J2, J3 = np.zeros((n, 3)), np.zeros((n, 9)), np.zeros((n, 9)), np.zeros((n, 18))
    for i in range(n):
        L.ref_motion_edge(P(p1[i]), P(p2[i]), P(mq[i]), P(mt[i]), float(dt[i]), P(mer[i]), P(J1[i]), P(J2[i]), P(J3[i]))
    out.update(motion_q=mq, motion_t=mt, motion_dt=dt, motion_err=mer, motion_J1=J1, motion_J2=J2, motion_J3=J3)
    mR = np.zeros((n, 9)); mupd = np.concatenate([rng.normal(0, 0.2, (n, 3)), rng.normal(0, 0.05, (n, 3))], 1)
    mupd[:4, 3:] = rng.normal(0, 0.8, (4, 3))                   # |v| > 1: fromCompactQuaternion returns the identity
    moR, mot = np.zeros((n, 9)), np.zeros((n, 3))
    for i in range(n):
        x, y, z, w = mq[i]
        mR[i] = [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]
        L.ref_motion_oplus(P(mR[i]), P(mt[i]), P(mupd[i]), P(moR[i]), P(mot[i]))
    out.update(motion_R=mR, motion_update=mupd, motion_oplus_R=moR, motion_oplus_t=mot)
    dst = os.path.join(ROOT, "tests", "golden", "ba_leaf_ref.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if not os.path.exists(LIB):
        sys.exit(f"{LIB} is missing: run `make -C oracle ref` in the container that holds /root/reference")
    main()
