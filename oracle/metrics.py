"""TEST INFRASTRUCTURE (CPU oracle) -- restatement of the reference's post-sampling evaluation metrics.

Follows, function by function (numpy, float32 inputs as the reference hands them over, float64 where it does):
  * compute_metrics_for_smpl      kinpoly/scripts/eval_metrics_imu_rec.py:264-342   (what eval_stage2.py:192 calls)
  * compute_foot_sliding_for_smpl kinpoly/scripts/eval_metrics_imu_rec.py:222-262
  * compute_accel                 kinpoly/scripts/eval_metrics_imu_rec.py:66-77
  * compute_error_accel           kinpoly/scripts/eval_metrics_imu_rec.py:79-107
  * get_root_matrix               kinpoly/relive/utils/metrics.py:15-24
  * get_frobenious_norm(_rot_only) kinpoly/relive/utils/metrics.py:64-82
  * quaternion_matrix             kinpoly/relive/utils/transformation.py:1346-1370  (wxyz, normalises, _EPS = 4 * eps)

Pinned by tests/golden/metrics.npz, produced by oracle/gen_golden_metrics.py which executes the reference's own function
bodies (extracted from the files above with `ast`; the modules themselves build a MuJoCo environment at import).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

EPS4 = np.finfo(float).eps * 4.0
H_ANKLE, H_TOE = 0.08, 0.04
FOOT_JOINTS = ((7, H_ANKLE), (10, H_TOE), (8, H_ANKLE), (11, H_TOE))     # l ankle, l toe, r ankle, r toe
HEAD_IDX = 15
# order of the vector the CUDA kernel returns (egoego_eval_metrics); single_jpe[22] follows
KEYS = ("root_trans_dist", "accel_pred", "accel_gt", "accel_err", "pred_fs", "gt_fs", "head_trans_dist", "root_dist",
        "root_rot_dist", "mpjpe", "mpjpe_wo_hand", "head_dist", "head_rot_dist")


def quaternion_matrix(q):
    q = np.array(q, dtype=np.float64, copy=True)
    n = np.dot(q, q)
    if n < EPS4:
        return np.identity(4)
    q *= np.sqrt(2.0 / n)
    q = np.outer(q, q)
    return np.array([[1.0 - q[2, 2] - q[3, 3], q[1, 2] - q[3, 0], q[1, 3] + q[2, 0], 0.0],
                     [q[1, 2] + q[3, 0], 1.0 - q[1, 1] - q[3, 3], q[2, 3] - q[1, 0], 0.0],
                     [q[1, 3] - q[2, 0], q[2, 3] + q[1, 0], 1.0 - q[1, 1] - q[2, 2], 0.0],
                     [0.0, 0.0, 0.0, 1.0]])


def pose_matrices(traj):
    out = []
    for pose in traj:
        m = quaternion_matrix(pose[3:7])
        m[:3, 3] = pose[:3]
        out.append(m)
    return out


def frobenius(x, y, rot_only=False):
    n = 3 if rot_only else 4
    err = 0.0
    for a, b in zip(x, y):
        err += np.linalg.norm(np.identity(n) - a[:n, :n] @ np.linalg.inv(b[:n, :n]), "fro")
    return err / len(x)


def compute_accel(joints):
    vel = joints[1:] - joints[:-1]
    acc = vel[1:] - vel[:-1]
    return np.mean(np.linalg.norm(acc, axis=2), axis=1)


def compute_error_accel(joints_gt, joints_pred):
    a_gt = joints_gt[:-2] - 2 * joints_gt[1:-1] + joints_gt[2:]
    a_pr = joints_pred[:-2] - 2 * joints_pred[1:-1] + joints_pred[2:]
    return np.mean(np.linalg.norm(a_pr - a_gt, axis=2), axis=1)


def foot_sliding(jpos, floor_height):
    jpos = jpos.copy()
    T = jpos.shape[0]
    jpos[:, :, 2] -= floor_height
    total = 0.0
    for j, h in FOOT_JOINTS:
        p = jpos[:, j, :]
        disp = np.linalg.norm(p[1:, :2] - p[:-1, :2], axis=1)
        sub = p[:-1, -1] < h
        total += np.sum(np.abs(disp * (2 - 2 ** (p[:-1, -1] / h)))[sub]) / T * 1000
    return total / 4.0


def compute_metrics_for_smpl(gt_quat, gt_jpos, gt_floor, pred_quat, pred_jpos, pred_floor):
    """Arrays [T,22,4] / [T,22,3] float32 -> dict with the reference's keys (plus single_jpe as a [22] vector)."""
    gt_quat, gt_jpos = np.asarray(gt_quat, np.float32), np.asarray(gt_jpos, np.float32)
    pred_quat, pred_jpos = np.asarray(pred_quat, np.float32), np.asarray(pred_jpos, np.float32)
    res = {}
    traj_p = np.concatenate((pred_jpos[:, 0], pred_quat[:, 0]), -1)
    traj_g = np.concatenate((gt_jpos[:, 0], gt_quat[:, 0]), -1)
    mp, mg = pose_matrices(traj_p), pose_matrices(traj_g)
    res["root_dist"], res["root_rot_dist"] = frobenius(mp, mg), frobenius(mp, mg, True)
    head_p = np.concatenate((pred_jpos[:, HEAD_IDX], pred_quat[:, HEAD_IDX]), -1)
    head_g = np.concatenate((gt_jpos[:, HEAD_IDX], gt_quat[:, HEAD_IDX]), -1)
    hp, hg = pose_matrices(head_p), pose_matrices(head_g)
    res["head_dist"], res["head_rot_dist"] = frobenius(hp, hg), frobenius(hp, hg, True)
    res["accel_pred"] = np.mean(compute_accel(pred_jpos)) * 1000
    res["accel_gt"] = np.mean(compute_accel(gt_jpos)) * 1000
    res["accel_err"] = np.mean(compute_error_accel(pred_jpos, gt_jpos)) * 1000
    res["pred_fs"], res["gt_fs"] = foot_sliding(pred_jpos, pred_floor), foot_sliding(gt_jpos, gt_floor)
    jp, jg = pred_jpos - pred_jpos[:, 0:1], gt_jpos - gt_jpos[:, 0:1]
    d = np.linalg.norm(jp - jg, axis=2)
    res["mpjpe"] = d.mean() * 1000
    single = d.mean(axis=0) * 1000
    res["mpjpe_wo_hand"] = single[:18].mean()
    res["root_trans_dist"] = np.linalg.norm(traj_p[:, :3] - traj_g[:, :3], axis=1).mean() * 1000
    res["head_trans_dist"] = np.linalg.norm(head_p[:, :3] - head_g[:, :3], axis=1).mean() * 1000
    res = {k: float(v) for k, v in res.items()}
    res["single_jpe"] = single.astype(np.float64)
    return res


def as_vector(res):
    """dict -> float64 [13 + 22] in the kernel's output order."""
    return np.concatenate((np.array([res[k] for k in KEYS], np.float64), np.asarray(res["single_jpe"], np.float64)))


def synth_motion(seed, T, J=22):
    """Seeded smooth-ish gt / pred joint trajectories and unit quaternions for metric tests: gt is a random walk of a
    random rest pose standing on z ~ 0, pred = gt + small perturbation, so every metric (incl. foot contacts) is exercised."""
    rng = np.random.default_rng(seed)
    rest = rng.uniform(-0.5, 0.5, (J, 3)).astype(np.float32)
    rest[:, 2] = rng.uniform(0.0, 1.7, J)
    rest[[7, 8], 2] = rng.uniform(0.02, 0.12, 2)          # ankles / toes near the floor thresholds
    rest[[10, 11], 2] = rng.uniform(0.0, 0.07, 2)
    walk = np.cumsum(rng.normal(0, 0.01, (T, 1, 3)), axis=0).astype(np.float32)
    wob = np.cumsum(rng.normal(0, 0.004, (T, J, 3)), axis=0).astype(np.float32)
    gt = rest[None] + walk + wob
    pred = gt + rng.normal(0, 0.02, (T, J, 3)).astype(np.float32) + np.cumsum(rng.normal(0, 0.002, (T, J, 3)), axis=0).astype(np.float32)
    gq = rng.normal(size=(T, J, 4)).astype(np.float32)
    gq /= np.linalg.norm(gq, axis=-1, keepdims=True)
    pq = gq + rng.normal(0, 0.05, (T, J, 4)).astype(np.float32)       # deliberately not unit: quaternion_matrix normalises
    return gq, gt.astype(np.float32), np.float32(rng.uniform(-0.01, 0.02)), pq.astype(np.float32), pred.astype(np.float32), np.float32(rng.uniform(-0.01, 0.02))
