"""Floor height / contacts: the numpy restatement against goldens produced by the reference's own function
with the real sklearn DBSCAN (oracle/gen_golden_floor.py), and the CUDA kernel (C ABI egoego_floor_contacts) against both."""
import os

import numpy as np
import pytest
import torch

from oracle import floor as OF
from oracle.gen_golden_floor import CASES


def _golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "floor.npz")))


def test_restatement_vs_reference_golden(golden_dir):
    g = _golden(golden_dir)
    for name, seed, T, kw in CASES:
        fh, contacts, discard = OF.determine_floor_height_and_contacts(OF.synth_walk(seed, T, **kw), 30)
        assert abs(float(fh) - float(g[f"{name}_floor"])) < 1e-7, name
        assert np.array_equal(contacts.astype(np.uint8), g[f"{name}_contacts"]), name
        assert bool(discard) == bool(g[f"{name}_discard"]), name
    assert bool(g["terrain_discard"]) and float(g["airborne_floor"]) == 0.0       # both special branches are in the goldens


def test_dbscan_restatement_vs_sklearn():
    """The 1-D DBSCAN restatement against scikit-learn itself (installed in this image) on random float32 data with clusters,
    chains, border points between clusters and noise -- labels compared as partitions plus the noise set."""
    sk = pytest.importorskip("sklearn.cluster")
    rng = np.random.default_rng(5)
    saw_noise = saw_multi = False
    for trial in range(60):
        n = int(rng.integers(1, 160))
        centers = rng.uniform(-0.05, 0.05, int(rng.integers(1, 6)))
        x = (rng.choice(centers, n) + rng.normal(0, rng.choice([0.0005, 0.002, 0.004]), n)).astype(np.float32)
        if trial % 3 == 0:
            x = np.round(x, 3).astype(np.float32)                   # many exact ties and gaps of exactly eps
        want = sk.DBSCAN(eps=0.005, min_samples=3).fit(x.reshape(-1, 1)).labels_
        got = OF.dbscan_1d(x)
        assert np.array_equal(got == -1, want == -1), trial
        pairs = {}
        for a, b in zip(got, want):
            assert pairs.setdefault(a, b) == b, trial               # same partition
        assert len(set(pairs.values())) == len(pairs), trial
        saw_noise |= bool((want == -1).any())
        saw_multi |= len(set(want[want >= 0])) > 1
    assert saw_noise and saw_multi


@pytest.mark.gpu
def test_gpu_floor_contacts_vs_reference_golden(golden_dir):
    import egoego_release_b200 as E
    g = _golden(golden_dir)
    for name, seed, T, kw in CASES:
        seq = torch.from_numpy(OF.synth_walk(seed, T, **kw)).cuda()
        fh, contacts, discard = E.determine_floor_height_and_contacts(seq, 30)
        assert abs(fh - float(g[f"{name}_floor"])) < 1e-6, (name, fh, float(g[f"{name}_floor"]))
        assert np.array_equal(contacts.astype(np.uint8), g[f"{name}_contacts"]), name
        assert discard == bool(g[f"{name}_discard"]), name
    with pytest.raises(E.EgoEgoError):
        E.determine_floor_height_and_contacts(torch.zeros(10, 22, 3), 30)          # CPU tensor: no fallback


@pytest.mark.gpu
def test_gpu_floor_contacts_batch_vs_oracle_random():
    """A batch of sequences in one launch, incl. heights rounded to the millimetre (ties, gaps of exactly eps, noise and border
    samples) and a long sequence (T = 700: 1400 samples through the O(n^2) passes), against the restatement."""
    import egoego_release_b200 as E
    for T, seeds in ((120, range(40, 56)), (700, range(60, 62))):
        seqs = []
        for s in seeds:
            q = OF.synth_walk(s, T, terrain=(s % 3 == 0))
            if s % 2 == 0:
                q[:, [OF.L_TOE, OF.R_TOE], 2] = np.round(q[:, [OF.L_TOE, OF.R_TOE], 2], 3)
            seqs.append(q)
        batch = torch.from_numpy(np.stack(seqs)).cuda()
        fl, ct, dc = E.floor_contacts_batch(batch, 30)
        for i, q in enumerate(seqs):
            fh, contacts, discard = OF.determine_floor_height_and_contacts(q, 30)
            assert abs(float(fl[i]) - float(fh)) < 1e-6, (T, i)
            assert np.array_equal(ct[i].cpu().numpy(), contacts.astype(np.float32)), (T, i)
            assert bool(dc[i]) == bool(discard), (T, i)
