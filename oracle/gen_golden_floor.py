"""Generate tests/golden/floor.npz by executing the UNMODIFIED reference functions ``determine_floor_height_and_contacts`` and
``detect_joint_contact`` (utils/data_utils/process_amass_dataset.py:160-328) with the real ``sklearn.cluster.DBSCAN``.

Run HERE only (needs /root/reference).  The module imports the licensed body model, matplotlib etc. at the top, so -- as for
the metrics goldens -- the function SOURCE is cut out with ``ast`` and exec'd unchanged in a namespace holding numpy, DBSCAN,
the module's constants (copied by name from its own assignments, :27-62) and SMPL_JOINTS (body_model/utils.py:5-8)."""
import ast
import os
import sys

import numpy as np
from sklearn.cluster import DBSCAN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
from oracle import floor as OF  # noqa: E402

CASES = [("walk", 31, 120, {}), ("walk_long", 32, 400, {}), ("terrain", 33, 240, {"terrain": True}),
         ("short", 34, 12, {}), ("airborne", 35, 60, {"airborne": True})]
CONSTS = ["FLOOR_VEL_THRESH", "FLOOR_HEIGHT_OFFSET", "CONTACT_VEL_THRESH", "CONTACT_TOE_HEIGHT_THRESH", "CONTACT_ANKLE_HEIGHT_THRESH",
          "TERRAIN_HEIGHT_THRESH", "ROOT_HEIGHT_THRESH", "CLUSTER_SIZE_THRESH", "DISCARD_TERRAIN_SEQUENCES", "VIZ_PLOTS"]


def reference_namespace():
    ns = {"np": np, "DBSCAN": DBSCAN}
    src = open(os.path.join(REF, "utils/data_utils/process_amass_dataset.py")).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and getattr(node.targets[0], "id", None) in CONSTS:
            exec(compile(ast.get_source_segment(src, node), "consts", "exec"), ns)
        if isinstance(node, ast.FunctionDef) and node.name in ("determine_floor_height_and_contacts", "detect_joint_contact"):
            exec(compile(ast.get_source_segment(src, node), node.name, "exec"), ns)
    usrc = open(os.path.join(REF, "body_model/utils.py")).read()
    for node in ast.parse(usrc).body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "SMPL_JOINTS":
            exec(compile(ast.get_source_segment(usrc, node), "SMPL_JOINTS", "exec"), ns)
    assert all(c in ns for c in CONSTS), [c for c in CONSTS if c not in ns]
    return ns


def main():
    ns = reference_namespace()
    out = {}
    for name, seed, T, kw in CASES:
        seq = OF.synth_walk(seed, T, **kw)
        fh, contacts, discard = ns["determine_floor_height_and_contacts"](seq, 30)
        out[f"{name}_floor"] = np.float64(fh)
        out[f"{name}_contacts"] = contacts.astype(np.uint8)
        out[f"{name}_discard"] = np.uint8(bool(discard))
        mf, mc, md = OF.determine_floor_height_and_contacts(seq, 30)
        print(f"{name}: floor {fh:.6f} (restatement {mf:.6f}), contacts sum {int(contacts.sum())} (equal: {np.array_equal(contacts, mc)}), "
              f"discard {discard}/{md}")
    np.savez(os.path.join(ROOT, "tests", "golden", "floor.npz"), **out)
    print("wrote tests/golden/floor.npz")


if __name__ == "__main__":
    main()
