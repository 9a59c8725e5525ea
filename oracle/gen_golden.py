"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run HERE only (the GPU box has no /root/reference); outputs are committed.

How the reference is made importable:
  * ``sys.modules`` stubs for packages that are absent from this image and only needed at import
    time: ``egoego.vis(.mesh_motion)``, ``body_model.body_model``,
    ``human_body_prior.body_model.body_model``.
  * ``pytorch3d.transforms`` (third-party, absent) is served by ``oracle/rotations.py``.  Goldens
    that flow through it (post-processing, sliding window) therefore pin the reference's *own*
    code (convert_model_res_to_data, quat_ik_torch, rotate_at_frame_smplh, the window loop), not
    pytorch3d: "parity unpinned" for those nine functions.
  * ``get_smpl_parents`` reads the licensed SMPL-H ``model.npz`` (absent); it is patched to return
    the standard 22-joint kintree.  ``ds`` is ``oracle.MotionDataStub`` (shipped stats + in-repo
    skeleton); its three methods restate AMASSDataset's (the class itself needs AMASS + CUDA).
  * Noise: ``torch.randn`` / ``torch.randn_like`` are patched with a numpy-Philox tape while the
    reference's own ``sample()`` / ``p_sample()`` / sliding-window loop run unmodified.
Weights come from ``oracle.init_params(seed)`` loaded with ``load_state_dict`` so tests can rebuild
them from the seed (44 MB of weights cannot be committed).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import egoego_oracle as O  # noqa: E402
from oracle import rotations as R  # noqa: E402


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dummy:  # BodyModel placeholder
        def __init__(self, *a, **k):
            raise RuntimeError("stub")

    p3d = stub("pytorch3d")
    tr = stub("pytorch3d.transforms", **{k: getattr(R, k) for k in dir(R) if not k.startswith("__")})
    p3d.transforms = tr
    stub("egoego.vis")
    stub("egoego.vis.mesh_motion", get_mesh_verts_faces_for_human_only=lambda *a, **k: None)
    stub("body_model")
    stub("body_model.body_model", BodyModel=_Dummy)
    stub("human_body_prior")
    stub("human_body_prior.body_model")
    stub("human_body_prior.body_model.body_model", BodyModel=_Dummy)
    sys.path.insert(0, REF)
    import egoego.model.transformer_cond_diffusion_model as M
    import egoego.data.amass_diffusion_dataset as DS
    parents = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19])
    DS.get_smpl_parents = lambda: parents.copy()
    M.quat_ik_torch.__globals__["get_smpl_parents"] = DS.get_smpl_parents
    return M


class Tape:
    """Patches torch.randn / randn_like with consecutive slices of a flat numpy-Philox stream."""

    def __init__(self, seed):
        self.rng = np.random.Generator(np.random.Philox(key=seed))

    def draw(self, shape):
        return torch.from_numpy(self.rng.standard_normal(size=tuple(shape), dtype=np.float32))

    def __enter__(self):
        self._r, self._rl = torch.randn, torch.randn_like
        torch.randn = lambda *s, **k: self.draw(s[0] if len(s) == 1 and not isinstance(s[0], int) else s)
        torch.randn_like = lambda x, **k: self.draw(x.shape)
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._r, self._rl


def build_model(M, params, timesteps, objective="pred_x0"):
    m = M.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256,
                                max_timesteps=121, out_dim=198, timesteps=timesteps,
                                objective=objective, loss_type="l1")
    missing, unexpected = m.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    assert all(k in O.make_schedule(10) for k in missing), missing  # only schedule buffers
    return m


def synth_x_start(seed, B, T):
    """Config-2 style conditioning: head channels only."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.zeros((B, T, 198), np.float32)
    x[:, :, 45:48] = rng.random((B, T, 3), dtype=np.float32) * 0.2 - 0.1
    yaw = (rng.random((B, T), dtype=np.float32) * 2 - 1) * np.float32(np.pi)
    tilt = (rng.random((B, T, 2), dtype=np.float32) * 2 - 1) * np.float32(0.3)
    aa = np.stack([tilt[..., 0], tilt[..., 1], yaw], -1)
    r6 = R.matrix_to_rotation_6d(R.axis_angle_to_matrix(torch.from_numpy(aa))).numpy()
    x[:, :, 156:162] = r6
    return torch.from_numpy(x)


def gen_pred_noise(M, params, out):
    """(vii) the constructor's default objective, `pred_noise` (transformer_cond_diffusion_model.py:233-236; no shipped script
    uses it): sample() at N=20, B=2, and three p_sample steps.  Kept in its own file so the other goldens stay byte-identical."""
    res = {}
    m = build_model(M, params, 20, objective="pred_noise")
    xs = synth_x_start(171, 2, 120)
    cm = O.prep_head_condition_mask(xs.shape)
    with Tape(71):
        res["sample_n20_b2_seed71"] = m.sample(xs, cm).numpy()
    m.eval()                                   # p_sample does not toggle eval itself (sample() does): dropout off
    rng = Tape(72)
    x, xc = rng.draw((2, 30, 198)), rng.draw((2, 30, 198))
    seq = []
    for t in (12, 7, 0):   # well-conditioned steps: at t = N-1 the cosine schedule has 1/abar ~ 1e5 and x0 amplifies fp32 rounding
        with rng:
            x = m.p_sample(x, torch.full((2,), t, dtype=torch.long), xc)
        seq.append(x.numpy().copy())
    res["p_sample_t30"] = np.stack(seq)
    np.savez(os.path.join(out, "pred_noise.npz"), **res)


def synth_head_pose(seed, B, T):
    """Seeded head trajectories [B,T,7] (xyz + wxyz): random-walk position around z = 1.6 m, yaw-dominant random-walk rotation."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = np.cumsum(rng.normal(0, 0.01, (B, T, 3)), axis=1).astype(np.float32) + np.float32([0, 0, 1.6])
    aa = np.stack([rng.normal(0, 0.05, (B, T)), rng.normal(0, 0.05, (B, T)), np.cumsum(rng.normal(0, 0.03, (B, T)), axis=1)], -1).astype(np.float32)
    quat = R.axis_angle_to_quaternion(torch.from_numpy(aa))
    return torch.cat((torch.from_numpy(pos), quat), -1)


def gen_sliding_edge(M, params, out):
    """(viii) window-length edge cases of p_sample_loop_sliding_window_w_canonical (:329-467), batch of 2 sequences, N = 20:
    121 frames (second window = the 10-frame overlap + ONE new frame, the shortest the loop produces) and 250 frames (three
    windows: 120, 120, 30)."""
    ds = O.MotionDataStub()
    m20 = build_model(M, params, 20)
    res = {}
    for T, seed in ((121, 61), (250, 62)):
        hp = synth_head_pose(seed, 2, T)
        data = torch.zeros(2, T, 198)
        cm = O.prep_head_condition_mask(data.shape)
        with Tape(seed):
            aa, root = m20.sample_sliding_window_w_canonical(ds, hp[:, :, :3], hp[:, :, 3:], x_start=data, cond_mask=cm)
        res[f"T{T}_aa"], res[f"T{T}_root"] = aa.numpy(), root.numpy()
        print("sliding edge", T, tuple(aa.shape))
    np.savez(os.path.join(out, "sliding_window_edge.npz"), **res)


def gen_sample_extra(M, params, out):
    """(ix) a second full-length golden: sample() of the unmodified reference at N = 1000 on FOUR windows with their own
    conditioning and noise (seed 23) -- the 1000-step parity claim does not rest on a single window."""
    N, B, seed = 1000, 4, 23
    m = build_model(M, params, N)
    xs = synth_x_start(2100, B, 120)
    cm = O.prep_head_condition_mask(xs.shape)
    with Tape(seed):
        y = m.sample(xs, cm)
    np.savez(os.path.join(out, "sample_extra.npz"), **{f"n{N}_b{B}_seed{seed}": y.numpy()})


def main():
    torch.set_num_threads(os.cpu_count())
    M = import_reference()
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    params = O.init_params(seed=0)
    ds = O.MotionDataStub()
    if "--only-sample-extra" in sys.argv:
        gen_sample_extra(M, params, out)
        print("wrote sample_extra.npz")
        return
    if "--only-sliding-edge" in sys.argv:
        gen_sliding_edge(M, params, out)
        return
    if "--only-pred-noise" in sys.argv:
        gen_pred_noise(M, params, out)
        print("wrote pred_noise.npz")
        return

    # (iv) schedule buffers, N=1000 and N=50
    for N in (1000, 50):
        m = build_model(M, params, N)
        np.savez(os.path.join(out, f"schedule_{N}.npz"),
                 **{k: v.numpy() for k, v in m.named_buffers() if "." not in k})

    m1000 = build_model(M, params, 1000).eval()

    # (i) denoiser forward
    fw = {}
    for tag, B, T, ts in (("b2_t120", 2, 120, (0, 999)), ("b1_t30", 1, 30, (500,))):
        x = torch.from_numpy(O.noise_tape(11, 1, (B, T, 396))[0])
        for t in ts:
            with torch.no_grad():
                y = m1000.denoise_fn(x, torch.full((B,), t, dtype=torch.long))
            fw[f"{tag}_t{t}"] = y.numpy()
    # padding-mask variant (rows zeroed after each sublayer, attention unmasked)
    x = torch.from_numpy(O.noise_tape(12, 1, (2, 120, 396))[0])
    pm = (torch.arange(121)[None, None, :] < torch.tensor([121, 61])[:, None, None])
    with torch.no_grad():
        fw["b2_t120_t7_padmask"] = m1000.denoise_fn(x, torch.full((2,), 7, dtype=torch.long), padding_mask=pm).numpy()
    np.savez(os.path.join(out, "denoiser_forward.npz"), **fw)

    # (ii) full sample(): N=50 B=2 and N=1000 B=1
    smp = {}
    for N, B, seed in ((50, 2, 21), (1000, 1, 22)):
        m = build_model(M, params, N)
        xs = synth_x_start(100 + N, B, 120)
        cm = O.prep_head_condition_mask(xs.shape)
        with Tape(seed):
            y = m.sample(xs, cm)
        smp[f"n{N}_b{B}_seed{seed}"] = y.numpy()
    np.savez(os.path.join(out, "sample.npz"), **smp)

    # (iii) p_sample + in-paint, 3 steps at t = 999, 998, 997 then t=1,0 (B=2, T=120 and T=30)
    ps = {}
    for T in (120, 30):
        rng = Tape(31 + T)
        x = rng.draw((2, T, 198))
        xc = rng.draw((2, T, 198))
        inp = rng.draw((2, 10, 198)).clamp(-1, 1)
        seq = []
        for t in (999, 998, 997, 1, 0):
            with rng:
                x = m1000.p_sample(x, torch.full((2,), t, dtype=torch.long), xc)
            x[:, :10, :] = inp
            seq.append(x.numpy().copy())
        ps[f"t{T}"] = np.stack(seq)
    np.savez(os.path.join(out, "p_sample_inpaint.npz"), **ps)

    # (v) post-processing on random + degenerate 6D inputs (reference convert_model_res_to_data)
    rng = Tape(41)
    xr = rng.draw((2, 16, 198)).clamp(-1, 1)
    xr[0, 0, 66:72] = 0.0                       # zero 6D
    xr[0, 1, 66:72] = torch.tensor([1., 0, 0, 2., 0, 0])   # collinear a1, a2
    xr[0, 2, 66:72] = torch.tensor([1e-9, 0, 0, 0, 1e-9, 0])  # tiny
    rec = O._np_normalize(rng.draw((2, 1, 1, 4)).numpy())
    aa, root, head = m1000.convert_model_res_to_data(ds, xr, rec.astype(np.float32), None)
    gq, gj = ds.fk_smpl(root.reshape(-1, 3), aa.reshape(-1, 22, 3))
    np.savez(os.path.join(out, "postprocess.npz"), aa=aa.numpy(), root=root.numpy(), head=head.numpy(),
             fk_jpos=gj.numpy(), fk_quat=gq.numpy())

    # (vi) sliding window on the 140-frame demo head pose, N=50 (windows 120 + 30)
    m50 = build_model(M, params, 50)
    hp = torch.from_numpy(np.load(os.path.join(out, "demo_head_qpos.npy")))[None]
    data = torch.zeros(1, hp.shape[1], 198)
    cm = O.prep_head_condition_mask(data.shape)
    with Tape(51):
        aa, root = m50.sample_sliding_window_w_canonical(ds, hp[:, :, :3], hp[:, :, 3:], x_start=data, cond_mask=cm)
    np.savez(os.path.join(out, "sliding_window.npz"), aa=aa.numpy(), root=root.numpy())
    gen_pred_noise(M, params, out)
    gen_sliding_edge(M, params, out)
    gen_sample_extra(M, params, out)
    print("goldens written:", sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
