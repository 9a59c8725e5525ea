"""Post-sampling evaluation metrics: oracle vs the reference's own functions (golden), and the
CUDA kernel (through the C ABI / host mirror) vs both.  Tolerance: 1e-4 relative (+1e-6 abs) -- the kernel accumulates
in fp64 like numpy but takes per-term norms in fp32 as the reference's float32 arrays do."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics as OM

CASES = [(11, 120), (12, 30), (13, 3), (14, 140)]
RTOL, ATOL = 1e-4, 1e-6


def _golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "metrics.npz")))


@pytest.mark.parametrize("seed,T", CASES)
def test_oracle_metrics_vs_reference_golden(seed, T, golden_dir):
    ref = _golden(golden_dir)[f"s{seed}_T{T}"]
    mine = OM.as_vector(OM.compute_metrics_for_smpl(*OM.synth_motion(seed, T)))
    np.testing.assert_allclose(mine, ref, rtol=1e-12, atol=1e-12)


def test_oracle_metrics_properties():
    gq, gj, gf, pq, pj, pf = OM.synth_motion(5, 40)
    same = OM.compute_metrics_for_smpl(gq, gj, gf, gq, gj, gf)
    for k in ("mpjpe", "root_trans_dist", "head_trans_dist", "accel_err", "root_dist", "head_rot_dist"):
        assert abs(same[k]) < 1e-6, k
    assert same["accel_pred"] == same["accel_gt"] and same["pred_fs"] == same["gt_fs"]
    # a rigid translation of both sequences changes nothing but the foot heights
    shifted = OM.compute_metrics_for_smpl(gq, gj + np.float32([1, 2, 0]), gf, pq, pj + np.float32([1, 2, 0]), pf)
    base = OM.compute_metrics_for_smpl(gq, gj, gf, pq, pj, pf)
    assert abs(shifted["mpjpe"] - base["mpjpe"]) < 1e-3 and abs(shifted["pred_fs"] - base["pred_fs"]) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("seed,T", CASES)
def test_gpu_metrics_vs_reference_golden(seed, T, golden_dir):
    from egoego_release_b200 import eval_metrics as EM
    ref = _golden(golden_dir)[f"s{seed}_T{T}"]
    gq, gj, gf, pq, pj, pf = [torch.as_tensor(a).cuda() for a in OM.synth_motion(seed, T)]
    res = EM.compute_metrics_for_smpl(gq, gj, float(gf), pq, pj, float(pf))
    got = np.array([res[k] for k in OM.KEYS] + [res["jpe_%d" % i] for i in range(22)])
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL)
    assert abs(res["single_jpe"] - ref[13:].mean()) <= RTOL * ref[13:].mean() + ATOL


@pytest.mark.gpu
def test_gpu_metrics_batch_and_errors():
    from egoego_release_b200 import eval_metrics as EM, EgoEgoError
    seqs = [OM.synth_motion(100 + i, 120) for i in range(5)]
    stack = lambda k: torch.as_tensor(np.stack([s[k] for s in seqs])).cuda()
    out = EM.compute_metrics_batch(stack(0), stack(1), stack(2), stack(3), stack(4), stack(5)).double().cpu().numpy()
    for i, s in enumerate(seqs):
        np.testing.assert_allclose(out[i], OM.as_vector(OM.compute_metrics_for_smpl(*s)), rtol=RTOL, atol=ATOL)
    # identical gt / pred: every distance is exactly zero
    z = EM.compute_metrics_batch(stack(0), stack(1), stack(2), stack(0), stack(1), stack(2)).cpu().numpy()
    assert np.all(z[:, [0, 3, 6, 9, 10]] == 0.0) and np.all(z[:, 13:] == 0.0)
    with pytest.raises(ValueError):
        EM.compute_metrics_batch(stack(0)[:, :2], stack(1)[:, :2], stack(2), stack(3)[:, :2], stack(4)[:, :2], stack(5))
    with pytest.raises(EgoEgoError):
        EM.compute_metrics_batch(stack(0).cpu(), stack(1).cpu(), stack(2).cpu(), stack(3).cpu(), stack(4).cpu(), stack(5).cpu())
