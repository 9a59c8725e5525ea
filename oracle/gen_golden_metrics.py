"""Generate tests/golden/metrics.npz by executing the UNMODIFIED reference metric functions.

Run HERE only (needs /root/reference).  kinpoly/scripts/eval_metrics_imu_rec.py builds a MuJoCo environment at import
time (mujoco_py, copycat: absent), so the module cannot be imported; instead the SOURCE of the functions on this path is
cut out of the reference files with `ast` and exec'd unchanged in a namespace that provides numpy / torch / math:

  kinpoly/scripts/eval_metrics_imu_rec.py : compute_accel, compute_error_accel, compute_foot_sliding_for_smpl,
                                            compute_metrics_for_smpl
  kinpoly/relive/utils/metrics.py         : get_root_matrix, get_frobenious_norm, get_frobenious_norm_rot_only
  kinpoly/relive/utils/transformation.py  : quaternion_matrix (+ its module constant _EPS)

Inputs are oracle.metrics.synth_motion(seed, T) (rebuilt from the seed by the tests); the committed file holds the
reference's outputs only.
"""
import ast
import math
import os
import sys
from collections import defaultdict

import numpy
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import metrics as OM  # noqa: E402

WANT = {
    "kinpoly/relive/utils/transformation.py": ["quaternion_matrix"],
    "kinpoly/relive/utils/metrics.py": ["get_root_matrix", "get_frobenious_norm", "get_frobenious_norm_rot_only"],
    "kinpoly/scripts/eval_metrics_imu_rec.py": ["compute_accel", "compute_error_accel", "compute_foot_sliding_for_smpl",
                                                "compute_metrics_for_smpl"],
}
CASES = [(11, 120), (12, 30), (13, 3), (14, 140)]


def reference_namespace():
    ns = {"np": np, "numpy": numpy, "math": math, "torch": torch, "defaultdict": defaultdict,
          "_EPS": numpy.finfo(float).eps * 4.0}           # transformation.py:1922 `_EPS = numpy.finfo(float).eps * 4.0`
    for rel, names in WANT.items():
        src = open(os.path.join(REF, rel)).read()
        tree = ast.parse(src)
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name in names:
                code = ast.get_source_segment(src, node)
                exec(compile(code, rel + ":" + node.name, "exec"), ns)
    return ns


def main():
    ns = reference_namespace()
    # make sure the constant really is what the reference defines
    tsrc = open(os.path.join(REF, "kinpoly/relive/utils/transformation.py")).read()
    assert "_EPS = numpy.finfo(float).eps * 4.0" in tsrc
    out = {}
    for seed, T in CASES:
        gq, gj, gf, pq, pj, pf = OM.synth_motion(seed, T)
        res = ns["compute_metrics_for_smpl"](torch.from_numpy(gq), torch.from_numpy(gj), gf,
                                             torch.from_numpy(pq), torch.from_numpy(pj), pf)
        # the reference collapses 'single_jpe' with np.mean; the per-joint values survive as jpe_<i>
        vec = np.array([res[k] for k in OM.KEYS] + [res["jpe_%d" % i] for i in range(22)], np.float64)
        out[f"s{seed}_T{T}"] = vec
        mine = OM.as_vector(OM.compute_metrics_for_smpl(gq, gj, gf, pq, pj, pf))
        print(f"seed {seed} T {T}: max rel diff restatement vs reference {np.max(np.abs(mine - vec) / (np.abs(vec) + 1e-12)):.2e}")
    np.savez(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **out)
    print("wrote tests/golden/metrics.npz")


if __name__ == "__main__":
    main()
