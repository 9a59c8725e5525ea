"""CPU restatement (numpy) of the floor-height / contact heuristic the evaluation scripts call before the metrics.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by the product.

Follows utils/data_utils/process_amass_dataset.py:160-328 (``determine_floor_height_and_contacts``, ``detect_joint_contact``;
constants :52-61; joint indices body_model/utils.py:5-8), called per sequence from eval_stage2.py:131,189 and
eval_egoego.py:331,395 (only the first return value, the offset floor height, is used there).

Third-party arithmetic: ``sklearn.cluster.DBSCAN(eps=0.005, min_samples=3)`` on the 1-D static toe heights
(requirements.txt lists scikit-learn without a version pin).  Its published algorithm is restated in ``dbscan_1d``:
  * neighbourhood = all points with |x_i - x_j| <= eps, the point itself included (distances in float64);
  * core point    = at least min_samples neighbours;
  * clusters are grown from core points in INDEX order (depth-first over core points); a non-core point within eps of a core
    point takes the label of the first cluster that reaches it; everything else is noise (label -1).
The reference then takes the median of EVERY label's heights -- the noise label included (np.unique(labels_)) -- and the floor is
the smallest median.  Pinned by tests/golden/floor.npz, produced by oracle/gen_golden_floor.py which exec's the reference's own
function bodies with the real sklearn DBSCAN.
"""
import numpy as np

HIPS, L_LEG, R_LEG, L_FOOT, R_FOOT, L_TOE, R_TOE, L_HAND, R_HAND = 0, 4, 5, 7, 8, 10, 11, 20, 21
FLOOR_VEL_THRESH = 0.005
FLOOR_HEIGHT_OFFSET = 0.01
CONTACT_VEL_THRESH = 0.005
CONTACT_TOE_HEIGHT_THRESH = 0.04
CONTACT_ANKLE_HEIGHT_THRESH = 0.08
TERRAIN_HEIGHT_THRESH = 0.04
ROOT_HEIGHT_THRESH = 0.04
CLUSTER_SIZE_THRESH = 0.25
DBSCAN_EPS, DBSCAN_MIN_SAMPLES = 0.005, 3


def dbscan_1d(x, eps=DBSCAN_EPS, min_samples=DBSCAN_MIN_SAMPLES):
    """Labels of sklearn's DBSCAN on 1-D data (see the module docstring)."""
    x = np.asarray(x, np.float64).reshape(-1)
    n = x.size
    nb = np.abs(x[:, None] - x[None, :]) <= eps
    core = nb.sum(1) >= min_samples
    labels = np.full(n, -1, np.int64)
    cur = 0
    for i in range(n):
        if labels[i] != -1 or not core[i]:
            continue
        stack = [i]
        while stack:                                  # sklearn/cluster/_dbscan_inner.pyx: depth-first growth over core points
            j = stack.pop()
            if labels[j] == -1:
                labels[j] = cur
                if core[j]:
                    for v in np.nonzero(nb[j])[0]:
                        if labels[v] == -1:
                            stack.append(v)
        cur += 1
    return labels


def _vel(seq):
    v = np.linalg.norm(seq[1:] - seq[:-1], axis=1)
    return np.append(v, v[-1])


def detect_joint_contact(seq, joint, floor_height, vel_thresh, height_thresh):
    js = seq[:, joint, :]
    return np.logical_and(_vel(js) < vel_thresh, js[:, 2] - floor_height < height_thresh)


def determine_floor_height_and_contacts(body_joint_seq, fps=30):
    """body_joint_seq [N, 22, 3] (z up) -> (offset_floor_height, contacts [N, 22], discard_seq); :160-317."""
    seq = np.asarray(body_joint_seq)
    n = seq.shape[0]
    root_h = seq[:, HIPS, 2]
    lt, rt = seq[:, L_TOE, :], seq[:, R_TOE, :]
    lv, rv = _vel(lt), _vel(rt)
    lh, rh = lt[:, 2], rt[:, 2]
    inds = np.arange(n)
    heights = np.append(lh[lv < FLOOR_VEL_THRESH], rh[rv < FLOOR_VEL_THRESH])
    sidx = np.append(inds[lv < FLOOR_VEL_THRESH], inds[rv < FLOOR_VEL_THRESH])
    discard = False
    if heights.shape[0] > 0:
        labels = dbscan_1d(heights)
        c_h, c_r, c_n = [], [], []
        min_median = min_root_median = float("inf")
        for lab in np.unique(labels):
            clust = heights[labels == lab]
            cinds = np.unique(sidx[labels == lab])
            med = np.median(clust)
            rmed = np.median(root_h[cinds])
            c_h.append(med); c_r.append(rmed); c_n.append(clust.shape[0])
            if med < min_median:
                min_median, min_root_median = med, rmed
        floor = min_median
        offset_floor = floor - FLOOR_HEIGHT_OFFSET
        for r, h, m in zip(c_r, c_h, c_n):
            if r > (min_root_median + ROOT_HEIGHT_THRESH) and h > (min_median + TERRAIN_HEIGHT_THRESH) and m > int(CLUSTER_SIZE_THRESH * fps):
                discard = True
                break
    else:
        floor = offset_floor = 0.0
    contacts = np.zeros((n, 22))
    lheel, rheel = seq[:, L_FOOT, :], seq[:, R_FOOT, :]
    contacts[:, L_FOOT] = np.logical_and(_vel(lheel) < CONTACT_VEL_THRESH, lheel[:, 2] - floor < CONTACT_ANKLE_HEIGHT_THRESH)
    contacts[:, R_FOOT] = np.logical_and(_vel(rheel) < CONTACT_VEL_THRESH, rheel[:, 2] - floor < CONTACT_ANKLE_HEIGHT_THRESH)
    contacts[:, L_TOE] = np.logical_and(lv < CONTACT_VEL_THRESH, lh - floor < CONTACT_TOE_HEIGHT_THRESH)
    contacts[:, R_TOE] = np.logical_and(rv < CONTACT_VEL_THRESH, rh - floor < CONTACT_TOE_HEIGHT_THRESH)
    for j in (L_HAND, R_HAND, L_LEG, R_LEG):
        contacts[:, j] = detect_joint_contact(seq, j, floor, CONTACT_VEL_THRESH, CONTACT_ANKLE_HEIGHT_THRESH)
    return offset_floor, contacts, discard


def synth_walk(seed, T, terrain=False, airborne=False):
    """Seeded [T, 22, 3] joint positions with alternating planted feet (static toes on the floor, a few on a step when
    ``terrain``), swing phases, slow drifts near the velocity threshold and jitter -- so the static sets, the DBSCAN
    clusters (several, plus noise points and border points between clusters) and every contact rule are exercised.
    ``airborne``: no toe is ever static (the reference's empty-set branch)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    j = rng.normal(0, 0.3, (T, 22, 3)).astype(np.float32)
    j[:, :, 2] += 1.0
    base = np.float32(rng.uniform(-0.05, 0.05))
    period = 24
    levels = np.float32([0.0, 0.009, 0.0, 0.021])                 # stance heights per step: clusters 9 / 12 mm apart (eps = 5 mm)
    raised = np.zeros(T, bool)
    for foot, toe, phase in ((L_FOOT, L_TOE, 0), (R_FOOT, R_TOE, period // 2)):
        pos = np.zeros((T, 3), np.float32)
        cur = np.array([0.0, 0.1 if toe == L_TOE else -0.1, base], np.float32)
        for t in range(T):
            k = (t + phase) % period
            if airborne or k >= period * 2 // 3:                  # swing: moves ~3 cm per frame and lifts
                cur = cur + np.float32([0.03, 0.0, 0.0])
                pos[t] = cur + np.float32([0, 0, 0.08 * np.sin(np.pi * (k - period * 2 // 3) / (period / 3))]) + (0.02 if airborne else 0.0)
            else:                                                  # stance: static up to sub-threshold drift / jitter
                step = ((t + phase) // period)
                lift = np.float32(0.12) if (terrain and step % 3 == 1) else np.float32(0.0)
                raised[t] |= bool(lift > 0)
                off = np.float32([0.0, 0.0, levels[step % 4]])
                if k == 5 and step % 2 == 0:                       # one stance frame per other step sits between the clusters:
                    off = off + np.float32([0, 0, 0.0045])         # a border / noise candidate for DBSCAN
                pos[t] = cur + off + np.float32([0, 0, lift]) + rng.normal(0, 0.0009, 3).astype(np.float32)
        j[:, toe, :] = pos
        j[:, foot, :] = pos + np.float32([-0.12, 0.0, 0.05]) + rng.normal(0, 0.0015, (T, 3)).astype(np.float32)
    j[:, HIPS, 2] = np.float32(0.9) + np.float32(0.12) * raised + rng.normal(0, 0.01, T).astype(np.float32)
    for jj in (L_HAND, R_HAND, L_LEG, R_LEG):                      # some frames of slow, low hands / knees
        lowmask = (np.arange(T) // 15) % 3 == 0
        j[lowmask, jj, :] = np.float32([0.3, 0.2, base + 0.03]) + rng.normal(0, 0.001, (int(lowmask.sum()), 3)).astype(np.float32)
    return j
