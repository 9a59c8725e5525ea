"""Generate tests/golden/stage1.npz by running the UNMODIFIED stage-1 reference modules from /root/reference.

Run HERE only.  How the modules are made importable:
  * pytorch3d.transforms (third-party, absent) is served by oracle/rotations.py -> "parity unpinned" for those functions;
  * evo (third-party, absent; only used by HeadNormalFormer.align_xy_plane_traj) is stubbed with an IDENTITY alignment, so
    HeadNormalFormer.forward_for_eval runs unmodified and its 'head_trans' / 'head_rot_mat' outputs pin everything it does
    before and after the xy-plane fit (rotation from the predicted floor normal, scale, re-integration of the trajectory);
    the Umeyama fit itself is NOT covered (not built either: DESIGN.md).
  * weights: oracle.stage1.init_params(seed, cfg) loaded with load_state_dict(strict=True); inputs:
    oracle.stage1.synth_stage1_inputs(seed, T).  The tests rebuild both from the seeds.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import rotations as R  # noqa: E402
from oracle import stage1 as S  # noqa: E402

HEAD_CASES = [(31, 139), (32, 45), (33, 60)]       # 3 blocks (60/60/19), one short block, exactly one window
NORMAL_CASES = [(41, 139), (42, 50), (43, 120)]    # cut to the window, padded, exactly window+1 poses


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    p3d = stub("pytorch3d")
    p3d.transforms = stub("pytorch3d.transforms", **{k: getattr(R, k) for k in dir(R) if not k.startswith("__")})

    class PoseTrajectory3D:       # identity stand-in for evo's trajectory + Umeyama alignment
        def __init__(self, positions_xyz=None, orientations_quat_wxyz=None, timestamps=None):
            self._positions_xyz = np.array(positions_xyz)

        def align(self, ref, correct_scale=False, correct_only_scale=False, n=-1):
            return np.eye(3), np.zeros(3), 1.0

    evo = stub("evo")
    core = stub("evo.core")
    evo.core = core
    core.trajectory = stub("evo.core.trajectory", PoseTrajectory3D=PoseTrajectory3D)
    core.sync = stub("evo.core.sync", associate_trajectories=lambda a, b: (a, b))
    sys.path.insert(0, REF)
    import egoego.model.head_estimation_transformer as HM
    import egoego.model.head_normal_estimation_transformer as NM
    return HM, NM


def main():
    HM, NM = import_reference()
    dev = torch.device("cpu")
    opt = argparse.Namespace(window=60, n_dec_layers=2, n_head=4, d_k=256, d_v=256, d_model=256, input_of_feats=True,
                             freeze_of_cnn=True, dist_scale=10.0, normal_window=120, normal_n_dec_layers=2, normal_n_head=4,
                             normal_d_k=256, normal_d_v=256, normal_d_model=256)
    out = {}
    with torch.no_grad():
        head = HM.HeadFormer(opt, dev)
        ph = S.init_params(7, S.CFG_HEAD)
        head.load_state_dict(ph, strict=True)
        head.eval()
        for seed, T in HEAD_CASES:
            feats, head_pose, slam_trans, slam_rot = S.synth_stage1_inputs(seed, T)
            data = {"of": feats, "aligned_slam_trans": slam_trans, "aligned_slam_rot_quat": R.matrix_to_quaternion(slam_rot),
                    "head_pose": head_pose}
            res = head.forward_for_eval(data)
            out[f"head_s{seed}_T{T}_pose"] = res["head_pose"].numpy()
            out[f"head_s{seed}_T{T}_scale"] = np.array(float(res["pred_scale"]))
            mine, sc = S.headformer_forward_for_eval(ph, feats, slam_trans, head_pose[:, 0, 3:])
            print(f"HeadFormer seed {seed} T {T}: restatement vs reference max-abs {float((mine - res['head_pose']).abs().max()):.2e}, "
                  f"scale {float(sc):.6f} vs {float(res['pred_scale']):.6f}")
        normal = NM.HeadNormalFormer(opt, dev, eval_whole_pipeline=True)
        pn = S.init_params(8, S.CFG_NORMAL)
        normal.load_state_dict(pn, strict=True)
        normal.eval()
        for seed, T in NORMAL_CASES:
            feats, head_pose, slam_trans, slam_rot = S.synth_stage1_inputs(seed, T)
            data = {"head_rot_mat": slam_rot, "head_trans": slam_trans, "seq_len": torch.tensor(slam_trans.shape[1]).float()[None],
                    "ori_head_pose": head_pose}
            normal.eval()          # forward_for_eval leaves action_transformer in train mode (:292): dropout would be live here
            fwd = normal.forward(data)["pred_normal"]
            scale = torch.tensor(2.5 + 0.1 * seed)
            ev = normal.forward_for_eval(data, pred_scale=scale)
            out[f"normal_s{seed}_T{T}_normal"] = fwd.numpy()
            out[f"normal_s{seed}_T{T}_trans"] = ev["head_trans"].numpy()          # R_xy = I: (trans_after - trans_after[0]) + gt[0]
            out[f"normal_s{seed}_T{T}_rot"] = ev["head_rot_mat"].numpy()
            mine = S.headnormal_forward(pn, slam_rot, slam_trans)
            ta, arm, _ = S.apply_normal_and_scale(mine, scale, slam_rot, slam_trans)
            ta = ta - ta[:, 0:1] + head_pose[:, 0:1, :3]
            print(f"HeadNormalFormer seed {seed} T {T}: normal max-abs {float((mine - fwd).abs().max()):.2e}, "
                  f"trans {float((ta - ev['head_trans']).abs().max()):.2e}, rot {float((arm - ev['head_rot_mat']).abs().max()):.2e}")
        # ResNet-18 optical-flow encoder: the reference's own wrapper class (egoego/model/resnet.py) around torchvision's resnet18
        from egoego.model.resnet import ResNet
        cnn = ResNet(512, running_stats=False, pretrained=False)
        pr = S.init_resnet_params(9)
        cnn.load_state_dict(pr, strict=True)
        cnn.eval()
        flow = S.synth_flow(51, 3)
        feats = cnn(S.flow_to_cnn_input(flow))
        out["resnet_s51_T3_feats"] = feats.numpy()
        print(f"ResNet-18: restatement vs reference max-abs {float((S.resnet18_forward(pr, S.flow_to_cnn_input(flow)) - feats).abs().max()):.2e} "
              f"(features in [{float(feats.min()):.3f}, {float(feats.max()):.3f}])")
    np.savez(os.path.join(ROOT, "tests", "golden", "stage1.npz"), **out)
    print("wrote tests/golden/stage1.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
