"""Derive the two small data assets the hot path needs from files the reference ships.

Run HERE only (needs /root/reference); the outputs are committed:

  egoego_release_b200/assets/smpl22_skeleton.json
      parents[22] + rest offsets[22,3] from the neutral SMPL body of
      kinpoly/assets/mujoco_models/humanoid_smpl_neutral_mesh.xml:48-176
      (``coordinate="global"`` body positions, :2) in the joint order of
      kinpoly/copycat/smpllib/smpl_parser.py:13-14.  Stand-in for the licensed
      SMPL-H ``model.npz`` the reference reads in
      egoego/data/amass_diffusion_dataset.py:83-90,248-263.
  egoego_release_b200/assets/cano_min_max_window_120.json
      global_jpos_min/max[22,3] from test_data/ares/cano_min_max_mean_std_data_window_120.p
      (the stats ``normalize_jpos_min_max`` uses, amass_diffusion_dataset.py:379-392).
  tests/golden/demo_head_qpos.npy
      head_qpos[140,7] of test_data/ares/demo_ares_data.p (real head-pose conditioning).
"""
import json
import os
import re
import sys

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NAMES = ['Pelvis', 'L_Hip', 'R_Hip', 'Torso', 'L_Knee', 'R_Knee', 'Spine', 'L_Ankle', 'R_Ankle',
         'Chest', 'L_Toe', 'R_Toe', 'Neck', 'L_Thorax', 'R_Thorax', 'Head', 'L_Shoulder',
         'R_Shoulder', 'L_Elbow', 'R_Elbow', 'L_Wrist', 'R_Wrist']


def main():
    xml = open(os.path.join(REF, "kinpoly/assets/mujoco_models/humanoid_smpl_neutral_mesh.xml")).read()
    # walk <body ...> / </body> to recover the tree
    pos, parent, stack = {}, {}, []
    for m in re.finditer(r'<body name="(\w+)" pos="([^"]+)"|</body>', xml):
        if m.group(0) == "</body>":
            stack.pop()
            continue
        name = m.group(1)
        pos[name] = [float(v) for v in m.group(2).split()]
        parent[name] = stack[-1] if stack else None
        stack.append(name)
    parents = [-1 if parent[n] is None else NAMES.index(parent[n]) for n in NAMES]
    assert parents == [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19], parents
    P = np.array([pos[n] for n in NAMES], dtype=np.float64)
    par0 = [0] + parents[1:]
    offsets = (P - P[par0]).astype(np.float32)
    out = {"joint_names": NAMES, "parents": parents,
           "rest_offsets": [[float(x) for x in r] for r in offsets],
           "source": "kinpoly/assets/mujoco_models/humanoid_smpl_neutral_mesh.xml (neutral SMPL, global body pos differences)"}
    os.makedirs(os.path.join(ROOT, "egoego_release_b200/assets"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "egoego_release_b200/assets/smpl22_skeleton.json"), "w"), indent=1)

    import joblib
    d = joblib.load(os.path.join(REF, "test_data/ares/cano_min_max_mean_std_data_window_120.p"))
    st = {"global_jpos_min": [float(x) for x in np.asarray(d["global_jpos_min"], np.float32)],
          "global_jpos_max": [float(x) for x in np.asarray(d["global_jpos_max"], np.float32)],
          "source": "test_data/ares/cano_min_max_mean_std_data_window_120.p"}
    json.dump(st, open(os.path.join(ROOT, "egoego_release_b200/assets/cano_min_max_window_120.json"), "w"), indent=1)

    e = joblib.load(os.path.join(REF, "test_data/ares/demo_ares_data.p"))
    os.makedirs(os.path.join(ROOT, "tests/golden"), exist_ok=True)
    np.save(os.path.join(ROOT, "tests/golden/demo_head_qpos.npy"), np.asarray(e[0]["head_qpos"], np.float32))
    print("ok", parents)


if __name__ == "__main__":
    sys.exit(main())
