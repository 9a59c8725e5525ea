"""Stand-in for the three ``AMASSDataset`` members the sampler touches (``self.ds`` in the reference
Trainer): normalisation statistics, the 22-joint kinematic tree / rest offsets and ``fk_smpl``
(egoego/data/amass_diffusion_dataset.py:248-293,379-392).  The real dataset needs licensed AMASS /
SMPL-H files; this object is built from the statistics the reference ships
(test_data/ares/cano_min_max_mean_std_data_window_120.p) and the neutral SMPL skeleton of
kinpoly/assets/mujoco_models/humanoid_smpl_neutral_mesh.xml (see oracle/make_assets.py).
A real ``AMASSDataset`` works equally: the engine only reads ``rest_human_offsets``,
``global_jpos_min/max`` and the parents table."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def smpl22_parents() -> np.ndarray:
    """The 22-joint SMPL kinematic tree (parents[0] = -1) shipped in assets/smpl22_skeleton.json; what the reference's
    ``get_smpl_parents()`` returns from the licensed SMPL-H ``model.npz`` (amass_diffusion_dataset.py:83-90)."""
    return np.asarray(json.load(open(os.path.join(ASSET_DIR, "smpl22_skeleton.json")))["parents"], dtype=np.int64)


class MotionDataStub:
    def __init__(self, device="cpu"):
        sk = json.load(open(os.path.join(ASSET_DIR, "smpl22_skeleton.json")))
        st = json.load(open(os.path.join(ASSET_DIR, "cano_min_max_window_120.json")))
        self.parents = np.asarray(sk["parents"], dtype=np.int64)
        self.rest_human_offsets = torch.tensor(sk["rest_offsets"], dtype=torch.float32, device=device)[None]
        self.global_jpos_min = torch.tensor(st["global_jpos_min"], dtype=torch.float32, device=device).reshape(22, 3)[None]
        self.global_jpos_max = torch.tensor(st["global_jpos_max"], dtype=torch.float32, device=device).reshape(22, 3)[None]
        self._engine = None

    def bind(self, model):
        """Route ``fk_smpl`` through ``model``'s CUDA engine (a CondGaussianDiffusion)."""
        self._engine = model
        return self

    def normalize_jpos_min_max(self, ori_jpos):
        mn, mx = self.global_jpos_min.to(ori_jpos.device), self.global_jpos_max.to(ori_jpos.device)
        return (ori_jpos - mn) / (mx - mn) * 2 - 1

    def de_normalize_jpos_min_max(self, normalized_jpos):
        mn, mx = self.global_jpos_min.to(normalized_jpos.device), self.global_jpos_max.to(normalized_jpos.device)
        return (normalized_jpos + 1) * 0.5 * (mx - mn) + mn

    def fk_smpl(self, root_trans, lrot_aa):
        if self._engine is None:
            raise RuntimeError("MotionDataStub.fk_smpl needs bind(model): FK runs in the CUDA library only")
        return self._engine.fk_smpl(self, root_trans, lrot_aa)
