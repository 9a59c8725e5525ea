"""oracle/rotations.py restates the nine pytorch3d.transforms functions of the path from their published
definitions (pytorch3d is third-party, absent from /root/reference and from this image: parity with pytorch3d ITSELF stays
unpinned).  What can be checked here is that the restatement computes the standard objects: every function is compared, in
float64, with scipy.spatial.transform.Rotation -- an independent implementation of the same conventions (Hamilton product,
active rotations; scipy stores quaternions scalar-LAST, pytorch3d scalar-first)."""
import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation as SR

from oracle import rotations as R


def _wxyz(q_xyzw):
    return np.concatenate((q_xyzw[..., 3:], q_xyzw[..., :3]), -1)


def _xyzw(q_wxyz):
    return np.concatenate((q_wxyz[..., 1:], q_wxyz[..., :1]), -1)


@pytest.fixture(scope="module")
def rots():
    rng = np.random.default_rng(5)
    rv = rng.normal(size=(500, 3)) * 1.2
    rv[0] = 0.0                                            # identity
    rv[1] = np.array([1e-9, 0, 0])                         # below the small-angle switch
    rv[2] = rv[2] / np.linalg.norm(rv[2]) * (np.pi - 1e-6)   # half turn: trace = -1 branch of matrix_to_quaternion
    rv[3] = np.array([0, np.pi, 0])
    return rv, SR.from_rotvec(rv)


def test_axis_angle_and_quaternion_to_matrix(rots):
    rv, sr = rots
    M = sr.as_matrix()
    assert np.abs(R.axis_angle_to_matrix(torch.from_numpy(rv)).numpy() - M).max() < 1e-12
    q = _wxyz(sr.as_quat())
    assert np.abs(R.quaternion_to_matrix(torch.from_numpy(q)).numpy() - M).max() < 1e-12
    # un-normalised quaternions describe the same rotation (the two_s = 2 / |q|^2 factor)
    assert np.abs(R.quaternion_to_matrix(torch.from_numpy(q * 3.7)).numpy() - M).max() < 1e-12
    qa = R.axis_angle_to_quaternion(torch.from_numpy(rv)).numpy()
    assert np.abs(np.linalg.norm(qa, axis=-1) - 1).max() < 1e-12 and np.abs(np.abs((qa * q).sum(-1)) - 1).max() < 1e-12


def test_matrix_to_quaternion_and_axis_angle(rots):
    rv, sr = rots
    M = torch.from_numpy(sr.as_matrix())
    q = R.matrix_to_quaternion(M).numpy()
    assert np.abs(np.linalg.norm(q, axis=-1) - 1).max() < 1e-12
    assert np.abs(np.abs((q * _wxyz(sr.as_quat())).sum(-1)) - 1).max() < 1e-12          # same rotation, sign free
    aa = R.matrix_to_axis_angle(M).numpy()
    assert np.abs(SR.from_rotvec(aa).as_matrix() - sr.as_matrix()).max() < 1e-9          # the half-turn case loses digits in acos/atan2
    qa = R.quaternion_to_axis_angle(torch.from_numpy(_wxyz(sr.as_quat()))).numpy()
    assert np.abs(SR.from_rotvec(qa).as_matrix() - sr.as_matrix()).max() < 1e-9


def test_quaternion_algebra(rots):
    rv, sr = rots
    a, b = sr[:250], sr[250:]
    qa, qb = torch.from_numpy(_wxyz(a.as_quat())), torch.from_numpy(_wxyz(b.as_quat()))
    prod = R.quaternion_multiply(qa, qb).numpy()
    ref = _wxyz((a * b).as_quat())
    assert (prod[:, 0] >= 0).all()                                                        # quaternion_multiply standardises the sign
    assert np.abs(prod - ref * np.where(ref[:, :1] < 0, -1.0, 1.0)).max() < 1e-12
    raw = R.quaternion_raw_multiply(qa, qb).numpy()
    assert np.abs(np.abs((raw * ref).sum(-1)) - 1).max() < 1e-12
    inv = R.quaternion_invert(qa).numpy()
    assert np.abs(SR.from_quat(_xyzw(inv)).as_matrix() - a.inv().as_matrix()).max() < 1e-12
    p = np.random.default_rng(6).normal(size=(250, 3))
    assert np.abs(R.quaternion_apply(qa, torch.from_numpy(p)).numpy() - a.apply(p)).max() < 1e-12
    with pytest.raises(ValueError):
        R.quaternion_apply(qa, torch.zeros(250, 4, dtype=torch.float64))


def test_rotation_6d_is_gram_schmidt():
    rng = np.random.default_rng(7)
    d6 = torch.from_numpy(rng.normal(size=(400, 6)))
    M = R.rotation_6d_to_matrix(d6).numpy()
    assert np.abs(M @ M.transpose(0, 2, 1) - np.eye(3)).max() < 1e-12 and np.abs(np.linalg.det(M) - 1).max() < 1e-12
    a1, a2 = d6[:, :3].numpy(), d6[:, 3:].numpy()
    b1 = a1 / np.linalg.norm(a1, axis=-1, keepdims=True)
    b2 = a2 - (b1 * a2).sum(-1, keepdims=True) * b1
    b2 /= np.linalg.norm(b2, axis=-1, keepdims=True)
    assert np.abs(M[:, 0] - b1).max() < 1e-12 and np.abs(M[:, 1] - b2).max() < 1e-12 and np.abs(M[:, 2] - np.cross(b1, b2)).max() < 1e-12
    # the 6D representation of a rotation matrix is its first two ROWS, and the map is idempotent on rotations
    back = R.matrix_to_rotation_6d(torch.from_numpy(M)).numpy()
    assert np.abs(back - np.concatenate((M[:, 0], M[:, 1]), -1)).max() == 0.0
    assert np.abs(R.rotation_6d_to_matrix(torch.from_numpy(back)).numpy() - M).max() < 1e-12
