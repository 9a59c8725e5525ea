"""The reference's REAL callers driving the mirror.

``/root/reference/trainer_amass_cond_motion_diffusion.py`` is imported unmodified (behind stubs for wandb, ema_pytorch,
pytorch3d -> oracle/rotations.py, the licensed body model and the visualisation modules) with ONE name swapped, exactly as
INTEGRATION.md documents: ``CondGaussianDiffusion`` -> the mirror class.  Its own ``Trainer`` is then constructed and its own
methods (``prep_head_condition_mask``, ``prep_padding_mask``, ``full_body_gen_cond_head_pose_sliding_window``, ``save`` / ``load``,
the EMA deep copy) call the mirror.

This container has no GPU, so the mirror class used here routes its five ENGINE calls (sampling loop, post-processing, FK,
canonicalisation, next-window conditioning) to the CPU oracle -- test infrastructure standing in for libegoego_b200 -- while every
line of the mirror's own host logic (sliding-window loop, stitching, eval/train toggles, state_dict handling) and of the reference
Trainer runs for real; the result must reproduce tests/golden/sliding_window.npz, which the unmodified reference model produced.
Skipped where /root/reference is absent (the GPU box); the same host logic is covered there by the ``-m gpu`` sliding-window tests.
"""
import copy
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (this container only)")

from oracle import egoego_oracle as O  # noqa: E402
from oracle import rotations as R  # noqa: E402
from oracle.gen_golden import Tape  # noqa: E402


class _EMA(torch.nn.Module):
    """Minimal stand-in for ema_pytorch.EMA (0.0.10): deep copy of the model as ``ema_model``, lerp update through ``.data``."""

    def __init__(self, model, beta=0.995, update_every=10, **kw):
        super().__init__()
        self.online_model = [model]                     # not registered (ema_pytorch keeps a plain reference too)
        self.ema_model = copy.deepcopy(model)
        self.ema_model.requires_grad_(False)
        self.beta = beta
        self.register_buffer("initted", torch.tensor(True))
        self.register_buffer("step", torch.tensor(0))

    def update(self):
        self.step += 1
        for pe, po in zip(self.ema_model.parameters(), self.online_model[0].parameters()):
            pe.data.lerp_(po.data, 1.0 - self.beta)


def _import_reference_trainer():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dummy:
        def __init__(self, *a, **k):
            raise RuntimeError("stub")

    p3d = stub("pytorch3d")
    p3d.transforms = stub("pytorch3d.transforms", **{k: getattr(R, k) for k in dir(R) if not k.startswith("__")})
    stub("wandb", init=lambda *a, **k: None, log=lambda *a, **k: None)
    stub("ema_pytorch", EMA=_EMA)
    stub("egoego.vis")
    stub("egoego.vis.mesh_motion", get_mesh_verts_faces_for_human_only=lambda *a, **k: None)
    stub("egoego.vis.blender_vis_mesh_motion", run_blender_rendering_and_save2video=lambda *a, **k: None,
         save_verts_faces_to_mesh_file=lambda *a, **k: None)
    stub("egoego.vis.pose", show3Dpose_animation_smpl22=lambda *a, **k: None)
    stub("body_model")
    stub("body_model.body_model", BodyModel=_Dummy)
    stub("human_body_prior")
    stub("human_body_prior.body_model")
    stub("human_body_prior.body_model.body_model", BodyModel=_Dummy)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import trainer_amass_cond_motion_diffusion as T
    return T


def _oracle_backed(E):
    """The mirror class with its ENGINE calls answered by the CPU oracle (see the module docstring)."""

    class OracleBacked(E.CondGaussianDiffusion):
        calls = []

        def _device(self):
            return torch.device("cpu")

        def _handle(self):
            return None

        def _params(self):
            return {k: v.detach() for k, v in self.state_dict().items()}

        def _sched(self):
            return {k: v for k, v in self.state_dict().items() if "." not in k}

        def p_sample_loop(self, shape, x_start, cond_mask, padding_mask=None, x_init=None, inpaint=None):
            OracleBacked.calls.append(("p_sample_loop", tuple(shape), inpaint is not None))
            tape, self._noise_tape = self._noise_tape, None
            assert tape is not None, "parity mode only: the caller must supply a noise tape"
            p, sched, N = self._params(), self._sched(), self.num_timesteps
            x = tape[0] if x_init is None else x_init.clone()
            x_cond = O.make_x_cond(x_start, cond_mask, tape[1])
            for k, i in enumerate(reversed(range(N))):
                x = O.p_sample(p, sched, x, i, x_cond, tape[2 + k], objective=self.objective)
                if inpaint is not None:
                    x[:, :inpaint.shape[1]] = inpaint
            return x

        def postprocess(self, ds, all_res_list, recover_rot_quat=None, with_fk=False):
            B = all_res_list.shape[0]
            rq = np.zeros((B, 1, 1, 4), np.float32)
            rq[..., 0] = 1
            if recover_rot_quat is not None:
                rq = np.asarray(torch.as_tensor(recover_rot_quat).reshape(B, 1, 1, 4).numpy(), np.float32)
            aa, root, head = O.convert_model_res_to_data(ds, all_res_list, rq)
            return (aa, root, head, None, None) if with_fk else (aa, root, head)

        def fk_smpl(self, ds, root_trans, lrot_aa):
            return ds.fk_smpl(root_trans, lrot_aa)

        def canonicalize_head(self, ds, head_jpos, head_jquat):          # reference :358-386
            b = head_jpos.shape[0]
            a_trans, a_quat, recover = O.rotate_at_frame_smplh(head_jpos.numpy(), head_jquat.numpy(), cano_t_idx=0)
            a_trans, a_quat = torch.from_numpy(a_trans), torch.from_numpy(a_quat)
            mz = a_trans[:, 0:1, :].clone()
            mz[:, :, 2] = 0
            a_trans = a_trans - mz
            a_r6 = R.matrix_to_rotation_6d(R.quaternion_to_matrix(a_quat))
            xs = torch.zeros(b, a_r6.shape[1], 198)
            xs[:, :, 45:48] = a_trans
            xs[:, :, 156:162] = a_r6
            xs[:, :, :66] = ds.normalize_jpos_min_max(xs[:, :, :66].reshape(-1, 22, 3)).reshape(b, -1, 66)
            return xs.float(), torch.from_numpy(np.asarray(recover, np.float32)).reshape(b, 4)

        def _tail_condition(self, ds, gq, gj):                           # reference :423-464
            b = gq.shape[0]
            t_trans, _, t_rec = O.rotate_at_frame_smplh(gj[:, :, 15, :].numpy(), gq[:, :, 15, :].numpy(), 0)
            tmz = torch.from_numpy(t_trans)[:, 0:1, :].clone()
            tmz[:, :, 2] *= 0
            inv = R.quaternion_invert(torch.from_numpy(t_rec).float()).repeat(1, gj.shape[1], gj.shape[2], 1)
            gj2 = R.quaternion_apply(inv, gj) - tmz[:, :, None, :]
            jp = ds.normalize_jpos_min_max(gj2.reshape(-1, 22, 3)).reshape(b, -1, 66)
            r6 = R.matrix_to_rotation_6d(R.quaternion_to_matrix(R.quaternion_multiply(inv, gq))).reshape(b, -1, 132)
            return torch.cat((jp, r6), dim=-1).float()

    return OracleBacked


class _FakeAMASS:
    """What Trainer.__init__ needs from AMASSDataset (amass_diffusion_dataset.py:146-): a map-style dataset of
    {'motion', 'seq_len'} plus the normalisation / FK members the sampler reads (the oracle's MotionDataStub provides them)."""

    def __init__(self, opt, train, window=120, run_demo=False):
        self._stub = O.MotionDataStub()
        for k in ("parents", "rest_human_offsets", "global_jpos_min", "global_jpos_max", "normalize_jpos_min_max",
                  "de_normalize_jpos_min_max", "fk_smpl"):
            setattr(self, k, getattr(self._stub, k))
        self.bm_dict = {}
        self.window = window

    def __len__(self):
        return 4

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(i)
        return {"motion": torch.rand(self.window, 198, generator=g) * 2 - 1, "seq_len": torch.tensor(60 + 20 * i)}


@pytest.fixture(scope="module")
def trainer_env(params0):
    import egoego_release_b200 as E
    T = _import_reference_trainer()
    T.CondGaussianDiffusion = _oracle_backed(E)        # the documented one-name swap
    T.AMASSDataset = _FakeAMASS
    opt = types.SimpleNamespace(window=120)
    model = T.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                    out_dim=198, timesteps=50, objective="pred_x0", loss_type="l1")
    model.load_state_dict(params0, strict=False)
    return T, T.Trainer(opt, model, train_batch_size=2, use_wandb=False, results_folder="/tmp/egoego_b200_trainer_test/weights"), E


def test_reference_trainer_constructs_with_the_mirror_class(trainer_env):
    T, trainer, E = trainer_env
    assert isinstance(trainer.model, E.CondGaussianDiffusion) and isinstance(trainer.ema.ema_model, E.CondGaussianDiffusion)
    assert trainer.ema.ema_model is not trainer.model and trainer.ema.ema_model._h is None       # EMA deep copy: no engine handle travels
    d = torch.zeros(2, 120, 198)
    assert torch.equal(trainer.prep_head_condition_mask(d), E.prep_head_condition_mask(d))        # reference helper == mirror's glue
    batch = next(trainer.dl)
    pm = trainer.prep_padding_mask(batch["motion"], batch["seq_len"])
    assert torch.equal(pm, E.prep_padding_mask(batch["motion"], batch["seq_len"]))
    # the checkpoint round trip of Trainer.save / Trainer.load (state_dict of model and EMA wrapper) through the mirror
    os.makedirs(trainer.results_folder, exist_ok=True)
    w0 = trainer.model.state_dict()["denoise_fn.linear_out.weight"].clone()
    trainer.save(1)
    with torch.no_grad():
        trainer.model.denoise_fn.linear_out.weight.mul_(0.0)
    trainer.load(1)
    assert torch.equal(trainer.model.state_dict()["denoise_fn.linear_out.weight"], w0)
    # ... and the 'ema' entry of that file loads into a bare mirror model, as INTEGRATION.md section 2 promises
    ck = torch.load(os.path.join(trainer.results_folder, "model-1.pt"))
    bare = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                   out_dim=198, timesteps=50, objective="pred_x0")
    bare.load_state_dict(ck["ema"], strict=False)
    assert torch.equal(bare.state_dict()["denoise_fn.linear_out.weight"], w0)


def test_reference_trainer_sliding_window_through_the_mirror(trainer_env, golden_dir):
    """Trainer.full_body_gen_cond_head_pose_sliding_window (reference :261-277) -> mirror.sample_sliding_window_w_canonical ->
    the mirror's own window loop / stitching; noise injected in the reference's draw order; result == the golden the unmodified
    reference model produced on the 140-frame demo head pose (two windows, 10-frame overlap in-painting)."""
    T, trainer, E = trainer_env
    g = dict(np.load(os.path.join(golden_dir, "sliding_window.npz")))
    hp = torch.from_numpy(np.load(os.path.join(golden_dir, "demo_head_qpos.npy")))[None]
    tape = Tape(51)
    mirror = trainer.ema.ema_model
    orig = mirror.sample_sliding_window_w_canonical
    mirror.sample_sliding_window_w_canonical = lambda *a, **k: orig(*a, noise_fn=tape.draw, **k)   # parity mode: the reference's draws
    type(mirror).calls.clear()
    try:
        aa, root = trainer.full_body_gen_cond_head_pose_sliding_window(hp, "demo")
    finally:
        del mirror.sample_sliding_window_w_canonical
    assert [c[0] for c in type(mirror).calls] == ["p_sample_loop", "p_sample_loop"] and type(mirror).calls[1][2]   # 2nd window in-paints
    assert tuple(aa.shape) == g["aa"].shape and tuple(root.shape) == g["root"].shape
    assert np.abs(root.numpy() - g["root"]).max() < 2e-4
    Ra, Rg = R.axis_angle_to_matrix(aa), R.axis_angle_to_matrix(torch.from_numpy(g["aa"]))
    assert float((Ra - Rg).abs().max()) < 2e-3
    assert mirror.denoise_fn.training                  # sample_* restores train() like the reference (:545)
