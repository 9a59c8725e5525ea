#!/usr/bin/env python
"""How ill-conditioned is the judged quantity (FK joint positions, mm) as a function of the raw sample, per weight set?
CPU only.  For each 1000-step golden of the unmodified reference (tests/golden/sample_ws_*.npz, sample_extra.npz): perturb the raw
normalised sample by uniform noise of amplitude 2.4e-6 (the raw max-abs distance between two fp32-grade implementations of the
sampler, profiles/r2e_parity_floor_*.txt) and report the resulting joint error, with the smallest 6D column norms and the largest
|cos| between the two 6D columns (rotation_6d_to_matrix normalises a1 and orthogonalises a2: errors grow like 1/|a1|, 1/sin)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import egoego_oracle as O

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ds = O.MotionDataStub()
for name, fn, key in (("seed0", "sample_extra.npz", "n1000_b4_seed23"), ("seed1", "sample_ws_seed1.npz", "n1000_b8"),
                      ("seed2", "sample_ws_seed2.npz", "n1000_b8"), ("trained_like", "sample_ws_trained_like.npz", "n1000_b8")):
    y = torch.from_numpy(np.load(os.path.join(G, fn))[key])
    j = O.joints_from_model_output(ds, y)
    g = torch.Generator().manual_seed(0)
    errs = []
    for _ in range(8):
        yp = y + 2.4e-6 * (torch.rand(y.shape, generator=g) * 2 - 1)
        errs.append(float((O.joints_from_model_output(ds, yp) - j).abs().max()) * 1e3)
    r6 = y[:, :, 66:].reshape(y.shape[0], y.shape[1], 22, 6)
    n1, n2 = r6[..., :3].norm(dim=-1), r6[..., 3:].norm(dim=-1)
    cs = ((r6[..., :3] * r6[..., 3:]).sum(-1) / (n1 * n2)).abs()
    print(f"{name:13s} joint max error for a 2.4e-6 raw perturbation: {min(errs):.3f} - {max(errs):.3f} mm over 8 draws | "
          f"min |a1| {float(n1.min()):.4f}, min |a2| {float(n2.min()):.4f}, max |cos(a1,a2)| {float(cs.max()):.6f}")
