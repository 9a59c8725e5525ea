#!/usr/bin/env python
"""bench.py -- motion-windows/sec of the stage-2 sampling path (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: a full N_diffusion(=1000)-step conditional sampling of
B (=256 per GPU) windows of T=120 frames (BASELINE config 2: "batch=256 T=120 1000-step sampling, random-init
denoiser, synthetic head-pose cond, 1xB200").  `value` is device-resident throughput (inputs already in HBM),
`e2e` goes through the host entry point (pinned host buffers, H2D + loop + D2H inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--diffusion-steps N]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_WINDOW_CALL = 2850.718e6      # algorithmic FLOPs per window per denoiser call (L=121)
# Per-kernel algorithmic work per window per launch (L = 121 real tokens, fp16 operand format bytes;
# weights are L2-resident and not counted).  DESIGN.md 5 carries the same table.
_ROW = 121 * 512 * 2                   # one fp16 activation row block [121, 512]
KERNEL_FLOPS = {"start": 24.330e6, "qkv": 380.633e6, "attention": 59.970e6, "fc_ln": 126.878e6, "w1": 63.439e6,
                "w2_ln": 63.439e6, "out": 24.330e6, "ddpm_update": 0.0}
KERNEL_BYTES = {"start": 120 * 198 * 2 + 121 * 512 * 4 + _ROW, "qkv": _ROW + 6 * _ROW, "attention": 6 * _ROW + 2 * _ROW,
                "fc_ln": 2 * _ROW + _ROW + _ROW, "w1": 2 * _ROW, "w2_ln": 3 * _ROW, "out": _ROW + 120 * 198 * 4,
                "ddpm_update": 3 * 120 * 198 * 4 + 120 * 198 * 2}
KERNEL_NAMES = {"start": "gemm_split3_2cta_kernel<TcEpiStart> (x half of start_conv)",
                "qkv": "gemm_half_tma_2cta_kernel<TmaEpiQKV> (fused QKV projection, TMA-store epilogue)",
                "attention": "attention_half_kernel (QK^T, softmax, PV; software-pipelined)",
                "fc_ln": "gemm_ln_half_c4_kernel (attention fc + residual + LayerNorm)",
                "w1": "gemm_half_tma_2cta_kernel<TmaEpiRelu> (FFN w_1 + ReLU)",
                "w2_ln": "gemm_ln_half_c4_kernel (FFN w_2 + residual + LayerNorm)",
                "out": "gemm_split3_2cta_kernel<TcEpiOutDdpm> (linear_out + fused DDPM update; TcEpiOut + ddpm_update_kernel if EGOEGO_FUSE_DDPM=0)",
                "ddpm_update": "ddpm_update_kernel"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="windows per GPU")
    ap.add_argument("--diffusion-steps", type=int, default=1000)
    ap.add_argument("--engine", default=None, choices=[None, "tcgen05", "simt"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline sample budget")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synth_inputs(B, T, seed=1234):
    import torch
    from oracle.gen_golden import synth_x_start
    from oracle import egoego_oracle as O
    xs = synth_x_start(seed, B, T)
    return xs, O.prep_head_condition_mask(xs.shape)


def cpu_baseline(B_cpu, T, N, budget_s):
    """The oracle port (torch CPU fp32 restatement of the reference sampler) on all host cores, bounded sample."""
    import torch
    from oracle import egoego_oracle as O
    cores = os.cpu_count() or 1
    params = O.init_params(0)
    sched = O.make_schedule(N)
    xs, cm = synth_inputs(B_cpu, T)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(xs.shape, generator=g)
    xc = O.make_x_cond(xs, cm, torch.randn(xs.shape, generator=g))
    with torch.no_grad():
        # give the CPU its best thread count (oversubscribing a 121-token problem on 100+ cores is slower)
        best, best_t = cores, None
        for nt in sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, 32), min(cores, 16)}, reverse=True):
            torch.set_num_threads(nt)
            O.p_sample(params, sched, x, N - 1, xc, torch.randn(xs.shape, generator=g))   # warm-up
            t0 = time.perf_counter()
            O.p_sample(params, sched, x, N - 1, xc, torch.randn(xs.shape, generator=g))
            dt = time.perf_counter() - t0
            if best_t is None or dt < best_t:
                best, best_t = nt, dt
        torch.set_num_threads(best)
        cores_used = best
        t0 = time.perf_counter(); n = 0
        while True:
            x = O.p_sample(params, sched, x, N - 1 - (n % N), xc, torch.randn(xs.shape, generator=g))
            n += 1
            el = time.perf_counter() - t0
            if (el > budget_s and n >= 3) or n >= N:
                break
    per_step = el / n
    return {"value": B_cpu / (per_step * N), "unit": "windows/s", "cores": cores_used, "host_cores": cores, "kind": "port",
            "sample": f"oracle p_sample, B={B_cpu}, T={T}, {n} of {N} steps timed ({el:.1f} s), scaled linearly to {N} steps",
            "ms_per_denoiser_step": per_step * 1e3}


def denoiser_latency(m, dev, T, iters=20):
    """BASELINE metric 2: device time of one denoiser forward (egoego_denoiser_forward, 3-term split format) in us."""
    import torch
    out = {}
    for B in (1, m._max_batch):
        x = torch.randn(B, T, 396, device=dev)
        t = torch.full((B,), m.num_timesteps // 2, dtype=torch.long, device=dev)     # 0 <= t < timesteps (the table has N rows)
        for _ in range(3):
            m.denoise_fn(x, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            m.denoise_fn(x, t)
        e1.record()
        torch.cuda.synchronize()
        out[f"B{B}"] = e0.elapsed_time(e1) / iters * 1e3
    out["note"] = "per call incl. staging of x/x_cond (egoego_denoiser_forward, fp16 hi/lo 3-term split), CUDA events, eager launches"
    return out


def parity_vs_reference(m, dev, N):
    """BASELINE metric 3: MPJPE / max-abs joint error (mm) of the product path against the committed golden output of the
    UNMODIFIED reference (tests/golden/sample.npz, produced by oracle/gen_golden.py) on identical conditioning and noise
    tape -- B=1, T=120, the bench's own 1000-step model and precision policy.  The oracle is only the checker here."""
    import numpy as np
    import torch
    from oracle import egoego_oracle as O
    from oracle.gen_golden import Tape, synth_x_start
    key = f"n{N}_b1_seed22"
    g = np.load(os.path.join(ROOT, "tests", "golden", "sample.npz"))
    if key not in g.files:
        return {"unavailable": f"no golden for {key}"}
    xs = synth_x_start(100 + N, 1, 120)
    cm = O.prep_head_condition_mask(xs.shape)
    tp = Tape(22)
    tape = torch.stack([tp.draw(xs.shape) for _ in range(N + 2)])
    m.set_noise_tape(tape.to(dev))
    y = m.sample(xs.to(dev), cm.to(dev)).cpu()
    m.set_noise_tape(None)
    ref = torch.from_numpy(g[key])
    ds = O.MotionDataStub()
    jy, jr = O.joints_from_model_output(ds, y), O.joints_from_model_output(ds, ref)
    return {"mpjpe_mm": O.mpjpe_mm(jy, jr), "joint_max_abs_mm": float((jy - jr).abs().max()) * 1e3,
            "raw_max_abs": float((y - ref).abs().max()), "tolerance_mm": 1.0,
            "reference": "unmodified reference CPU fp32 (golden tests/golden/sample.npz), B=1 T=120 N=%d, identical noise tape" % N}


def next_rows(dev, B, T, pk):
    """SURVEY.md 8f rows built beside the hot path, each measured on the device with CUDA events:
    the evaluation-metrics kernel (HBM-bound: algorithmic bytes = joint positions + root/head quaternions of gt and pred) and the stage-1
    networks (latency: one 139-frame sequence, the demo's length, through HeadFormer / HeadNormalFormer forward_for_eval)."""
    import argparse
    import numpy as np
    import torch
    import egoego_release_b200 as E
    from oracle import stage1 as S
    out = {}

    def timed(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    nseq = 8 * B                                     # 2048 sequences: 300 MB of inputs, larger than the 126 MB L2
    gq = torch.randn(nseq, T, 22, 4, device=dev); gj = torch.randn(nseq, T, 22, 3, device=dev)
    pq = torch.randn(nseq, T, 22, 4, device=dev); pj = gj + 0.01 * torch.randn(nseq, T, 22, 3, device=dev)
    fl = torch.zeros(nseq, device=dev)
    ms = timed(lambda: E.compute_metrics_batch(gq, gj, fl, pq, pj, fl), 10)
    by = nseq * T * (22 * 3 + 2 * 4) * 4 * 2         # positions of 22 joints + quaternions of root and head, gt and pred
    out["eval_metrics"] = {"kernel": "eval_metrics_kernel (compute_metrics_for_smpl)", "sequences": nseq, "ms_per_launch": ms, "bound": "hbm",
                           "algorithmic_bytes_per_launch": by, "achieved": by / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                           "frac": by / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "sequences_per_s": nseq / (ms * 1e-3)}
    del gq, gj, pq, pj
    opt = argparse.Namespace(window=60, n_dec_layers=2, n_head=4, d_k=256, d_v=256, d_model=256, input_of_feats=True, freeze_of_cnn=True,
                             dist_scale=10.0, normal_window=120, normal_n_dec_layers=2, normal_n_head=4, normal_d_k=256, normal_d_v=256,
                             normal_d_model=256)
    hf = E.HeadFormer(opt, dev); hf.load_state_dict(S.init_params(7, S.CFG_HEAD)); hf = hf.to(dev)
    gn = E.HeadNormalFormer(opt, dev, eval_whole_pipeline=True); gn.load_state_dict(S.init_params(8, S.CFG_NORMAL)); gn = gn.to(dev)
    feats, head_pose, slam_trans, slam_rot = [t.to(dev) for t in S.synth_stage1_inputs(77, 139)]
    d1 = {"of": feats, "aligned_slam_trans": slam_trans, "head_pose": head_pose}
    d2 = {"head_rot_mat": slam_rot, "head_trans": slam_trans, "ori_head_pose": head_pose}
    sc = torch.tensor(1.0, device=dev)
    ident = lambda a, b: np.eye(3)
    l0 = hf.launch_count()
    t1 = timed(lambda: hf.forward_for_eval(d1), 10)
    out["stage1"] = {"sequence_frames": 139, "headformer_forward_for_eval_us": t1 * 1e3,
                     "headnormalformer_forward_us": timed(lambda: gn.forward(d2), 10) * 1e3,
                     "headnormalformer_forward_for_eval_us": timed(lambda: gn.forward_for_eval(d2, sc, xy_align=ident), 10) * 1e3,
                     "note": "fp32 CUDA-core sequence nets (d_model 256, 2 layers), eager launches incl. host glue; forward_for_eval of "
                             "HeadNormalFormer includes its device->host copy of the trajectory for the xy alignment callable",
                     "headformer_launches_per_call": (hf.launch_count() - l0) // 13}
    # the same two networks as stock PyTorch on this GPU (oracle port of the reference modules' op sequence, fp32 eager)
    try:
        ph = {k: v.to(dev) for k, v in S.init_params(7, S.CFG_HEAD).items()}
        pn = {k: v.to(dev) for k, v in S.init_params(8, S.CFG_NORMAL).items()}
        q0 = head_pose[:, 0, 3:]
        ot = slam_trans - slam_trans[:, 0:1]
        with torch.no_grad():
            out["stage1"]["torch_eager_headformer_forward_for_eval_us"] = timed(lambda: S.headformer_forward_for_eval(ph, feats, slam_trans, q0), 10) * 1e3
            out["stage1"]["torch_eager_headnormalformer_forward_us"] = timed(lambda: S.headnormal_forward(pn, slam_rot, ot), 10) * 1e3
    except Exception as ex:
        out["stage1"]["torch_eager_error"] = repr(ex)[:200]
    # HeadNet with raw optical flow: ResNet-18 encoder (1.814 GFLOP per 224 x 224 frame, multiply-add = 2) on the demo's 139 frames
    try:
        opt2 = argparse.Namespace(**{**vars(opt), "input_of_feats": False})
        hf2 = E.HeadFormer(opt2, dev)
        hf2.load_state_dict({**S.init_params(7, S.CFG_HEAD), **{"cnn." + k: v for k, v in S.init_resnet_params(9).items()}})
        hf2 = hf2.to(dev)
        flow = torch.randn(1, 139, 224, 224, 2, device=dev)
        ms = timed(lambda: hf2._input_features({"of": flow}), 5)
        out["resnet18_encoder"] = {"frames": 139, "ms_per_call": ms, "frames_per_s": 139 / (ms * 1e-3),
                                   "achieved": 139 * 1.814e9 / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                   "note": "convolutions with folded BatchNorm and fused bias + residual + ReLU epilogues (csrc/resnet.cu)"}
        # the reference's ResNet class as stock PyTorch on this GPU: cuDNN convolutions, fp32 with and without TF32 (torch's default
        # for convolutions is allow_tf32 = True -- the reference's real numerics on a GPU)
        pr = {k: v.to(dev) for k, v in S.init_resnet_params(9).items()}
        xin = S.flow_to_cnn_input(flow.cpu()).to(dev)
        for tag, tf32 in (("cudnn_tf32_default", True), ("cudnn_fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                mt = timed(lambda: S.resnet18_forward(pr, xin), 5)
            out["resnet18_encoder"]["torch_" + tag] = {"ms_per_call": mt, "frames_per_s": 139 / (mt * 1e-3)}
        torch.backends.cudnn.allow_tf32 = True
    except Exception as ex:
        out["resnet18_encoder"] = {"error": repr(ex)[:200]}
    # training step (BASELINE configs[4] shape per GPU: batch 32, T = 120, train() mode with dropout): the reference Trainer's
    # loop body -- autocast(fp16) + GradScaler + Adam -- through egoego_train_step vs the reference's op sequence as stock PyTorch
    # under the same autocast on this GPU (tools/train_bench.py; the 8-GPU DDP run of the same tool is profiles/r2*_train_ddp*.json)
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_bench
        out["train_step"] = train_bench.run_arms(dev, 32, T, steps=10, warmup=3)
    except Exception as ex:
        out["train_step"] = {"error": repr(ex)[:300]}
    return out


def config1_latency(dev, T):
    """BASELINE configs[0] shape on the GPU: ONE window, T=120, a 50-step schedule (the reference's CPU-runnable case; the
    survey probe measured 0.68 s on 8 CPU cores) -- latency of sample() through the host mirror, CUDA events, plus B=1 at the
    full 1000-step schedule."""
    import torch
    import egoego_release_b200 as E
    from oracle import egoego_oracle as O
    out = {}
    # (B, N): one window at the 50-step and the full schedule; B = 32 = --diffusion_batch_size of scripts/test_egoego_pipeline.sh
    for Bw, N in ((1, 50), (1, 1000), (32, 1000)):
        xs, cm = synth_inputs(Bw, T)
        xs, cm = xs.to(dev), cm.to(dev)
        m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                    out_dim=198, timesteps=N, objective="pred_x0", loss_type="l1", max_batch=Bw)
        m.load_state_dict(O.init_params(0), strict=False)
        m = m.to(dev)
        for _ in range(2):
            m.sample(xs, cm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5 if N == 50 else 2
        for _ in range(reps):
            m.sample(xs, cm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[f"B{Bw}_T{T}_N{N}"] = {"ms_per_sample_call": ms, "windows_per_s": Bw * 1e3 / ms, "us_per_diffusion_step": ms * 1e3 / N,
                                   "precise_last_steps": m.precise_last_steps()}
        del m
    return out


def all_split_ms(dev, xs, cm, B, T, N):
    """One sample() of the bench workload with precise_last_steps = N (all 3-term split), CUDA events."""
    import torch
    import egoego_release_b200 as E
    from oracle import egoego_oracle as O
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                out_dim=198, timesteps=N, objective="pred_x0", loss_type="l1", max_batch=B, precise_last_steps=N)
    m.load_state_dict(O.init_params(0), strict=False)
    m = m.to(dev)
    m.sample(xs, cm)                                           # first call of THIS handle: weight commit, graph capture
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.sample(xs, cm)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def pipeline_config4(dev):
    """BASELINE configs[3]: the full EgoEgo pipeline on one B200 -- HeadNet + GravityNet on a 139-frame sequence (the demo's
    length) -> stage-2 sliding-window diffusion (2 windows x 1000 steps, sample batch 32 = scripts/test_egoego_pipeline.sh) ->
    FK -> floor height + evaluation metrics, end to end through pipeline.run_egoego; CUDA events around whole calls."""
    import argparse
    import numpy as np
    import torch
    import egoego_release_b200 as E
    from egoego_release_b200 import pipeline as P
    from oracle import egoego_oracle as O
    from oracle import stage1 as S
    out = {}
    opt = argparse.Namespace(window=60, n_dec_layers=2, n_head=4, d_k=256, d_v=256, d_model=256, input_of_feats=True, freeze_of_cnn=True,
                             dist_scale=10.0, normal_window=120, normal_n_dec_layers=2, normal_n_head=4, normal_d_k=256, normal_d_v=256,
                             normal_d_model=256)
    hf = E.HeadFormer(opt, dev); hf.load_state_dict(S.init_params(7, S.CFG_HEAD)); hf = hf.to(dev)
    gn = E.HeadNormalFormer(opt, dev, eval_whole_pipeline=True); gn.load_state_dict(S.init_params(8, S.CFG_NORMAL)); gn = gn.to(dev)
    feats, head_pose, slam_trans, slam_rot = S.synth_stage1_inputs(77, 139)
    data = {"of": feats, "aligned_slam_trans": slam_trans, "head_pose": head_pose, "ori_slam_trans": slam_trans * 1.9, "ori_slam_rot_mat": slam_rot}
    ident = lambda est, ref: np.eye(3)
    for bs in (1, 32):
        dm = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                     out_dim=198, timesteps=1000, objective="pred_x0", loss_type="l1", max_batch=bs)
        dm.load_state_dict(O.init_params(0), strict=False)
        dm = dm.to(dev)
        ds = E.MotionDataStub().bind(dm)

        def run():
            o = P.run_egoego(hf, gn, dm, ds, data, sample_bs=bs, xy_align=ident)
            fl, _, _ = E.floor_contacts_batch(o["global_jpos"], 30)
            return E.compute_metrics_batch(o["global_jrot"], o["global_jpos"], fl, o["global_jrot"], o["global_jpos"], fl)

        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[f"sample_bs{bs}"] = {"ms_per_sequence_call": ms, "sequences_per_s": bs * 1e3 / ms, "frames": 139, "windows_per_sequence": 2,
                                 "diffusion_steps_per_window": 1000, "finite": bool(torch.isfinite(res).all())}
        del dm
    out["note"] = ("stage 1 (HeadFormer + HeadNormalFormer forward_for_eval) -> sliding-window stage 2 -> FK -> floor height -> metrics, all on "
                   "the device; the two windows of a sequence are sequentially dependent (10-frame overlap in-painting), so the time is "
                   "2 x 1000 diffusion steps at batch = sample_bs")
    return out


def torch_gpu_baseline(dev, B, T, N, n_steps=12):
    """The reference's own arithmetic (oracle port: same op sequence) as stock PyTorch eager fp32 on this B200 --
    the denominator of north_star's ">= 10x the reference single-GPU PyTorch sampling throughput"."""
    import torch
    from oracle import egoego_oracle as O
    params = {k: v.to(dev) for k, v in O.init_params(0).items()}
    sched = {k: v.to(dev) for k, v in O.make_schedule(N).items()}
    xs, cm = synth_inputs(B, T)
    xs, cm = xs.to(dev), cm.to(dev)
    out = {}
    for tag, conv in (("linear_fp32", False), ("conv1d_cudnn_tf32_default", True)):
        O.USE_CONV1D = conv
        x = torch.randn(xs.shape, device=dev)
        xc = O.make_x_cond(xs, cm, torch.randn(xs.shape, device=dev))
        with torch.no_grad():
            for i in range(3):
                x = O.p_sample(params, sched, x, N - 1 - i, xc, torch.randn(xs.shape, device=dev))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n_steps):
                x = O.p_sample(params, sched, x, N - 4 - i, xc, torch.randn(xs.shape, device=dev))
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_steps
        out[tag] = {"windows_per_s": B / (ms * 1e-3 * N), "ms_per_denoiser_step": ms}
    O.USE_CONV1D = False
    out["kind"] = (f"oracle port (reference op sequence) in PyTorch eager on the same GPU, B={B}, {n_steps} steps timed with CUDA events, "
                   "scaled to 1000 steps; conv1d variant = reference's Conv1d k=1 layers via cuDNN with stock TF32 flags")
    return out


def main():
    a = parse()
    os.environ.setdefault("TQDM_DISABLE", "1")
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    T, D, N, B = 120, 198, a.diffusion_steps, a.batch
    cfg = {"workload": f"configs[1]: batch={B}/GPU T={T} {N}-step sampling, random-init denoiser (oracle.init_params seed 0), "
                       "synthetic head-pose cond", "windows_per_gpu": B, "T": T, "diffusion_steps": N, "d_feats": D,
           "parallelism": (f"configs[2]: windows sharded over {world} GPU(s), per-rank post-processing, ONE all-gather of the generated SMPL "
                           "parameters (69 floats per frame)") if world > 1 else "single GPU",
           "l2": "per-step working set (fp16/fp32 activation planes for 256 windows, >1 GB) exceeds the 126 MB L2; no flush"}

    if a.impl == "reference":
        if rank != 0:
            return
        cb = cpu_baseline(32, T, N, max(a.cpu_seconds, 10.0) * max(1, a.steps) / 3.0)
        line = {"impl": "reference", "metric": "motion-windows/sec (T=120, 1000-step)", "value": cb["value"], "unit": "windows/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": B / cb["value"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "reference arm = CPU oracle port of the reference sampler on the host cores (the Python reference "
                        "itself cannot travel to the GPU box); bounded sample scaled to the full workload"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import egoego_release_b200 as E
    from oracle import egoego_oracle as O
    m = E.CondGaussianDiffusion(d_feats=D, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                out_dim=D, timesteps=N, objective="pred_x0", loss_type="l1", max_batch=B, engine=a.engine)
    m.load_state_dict(O.init_params(0), strict=False)
    m = m.to(dev)
    m.window_offset = rank * B                      # Philox streams keyed by GLOBAL window id
    xs_h, cm_h = synth_inputs(B * world, T)
    xs_h = xs_h[rank * B:(rank + 1) * B].contiguous().pin_memory()
    cm_h = cm_h[rank * B:(rank + 1) * B].contiguous().pin_memory()
    out_h = torch.empty(B, T, D).pin_memory()
    xs, cm = xs_h.to(dev), cm_h.to(dev)
    # configs[2]: the ONE collective of the path gathers the generated SMPL parameters (22 local axis-angles + root translation =
    # 69 floats per frame, 8.5 MB per rank at 256 windows) -- each rank post-processes its own windows first (parallel.smpl_post_fn)
    from egoego_release_b200 import parallel as PAR
    ds = E.MotionDataStub().bind(m)
    post = PAR.smpl_post_fn(m, ds)
    gathered = torch.empty(world * B, T, 69, device=dev) if world > 1 else None
    torch.manual_seed(1234)

    def step_device():
        y = m.sample(xs, cm)
        if world > 1:
            dist.all_gather_into_tensor(gathered, post(y, rank * B).contiguous())
        return y

    def step_host():
        y = m.sample_host(xs_h, cm_h, out=out_h)
        if world > 1:
            dist.all_gather_into_tensor(gathered, post(y.to(dev, non_blocking=True), rank * B).contiguous())
        return y

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = m.launch_count()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms, m.launch_count() - l0

    for _ in range(max(a.warmup, 3)):
        step_device()
    with ClockSampler(local) as cs:
        ms, launches = timed(step_device, a.steps)
    clocks = cs.summary()
    step_host()                                          # warm the host path (staging buffers)
    ms_h, _ = timed(step_host, a.steps)

    # shard invariance, checked on the device at every N > 1 ("8-GPU output == 1-GPU output window for window"):
    # with a fixed seed, rank 0 re-samples the LAST rank's shard (same conditioning, Philox streams keyed by the global window id)
    # and compares its SMPL parameters bit for bit with what the all-gather delivered from that rank.
    shard_check = None
    if world > 1:
        torch.manual_seed(777)                           # every rank draws the same sampling seed from torch's generator
        y = m.sample(xs, cm)
        dist.all_gather_into_tensor(gathered, post(y, rank * B).contiguous())
        xs_all, cm_all = synth_inputs(B * world, T)
        if rank == 0:
            r = world - 1
            m.window_offset = r * B
            torch.manual_seed(777)
            y_r = m.sample(xs_all[r * B:(r + 1) * B].to(dev), cm_all[r * B:(r + 1) * B].to(dev))
            m.window_offset = 0
            mine = post(y_r, r * B)
            theirs = gathered[r * B:(r + 1) * B]
            import hashlib
            shard_check = {"rank_checked": r, "bit_identical": bool(torch.equal(mine, theirs)),
                           "max_abs_diff": float((mine - theirs).abs().max()),
                           "sha256_gathered_smpl_params": hashlib.sha256(gathered.cpu().numpy().tobytes()).hexdigest()[:16],
                           "gathered_bytes_per_rank": int(B * T * 69 * 4)}
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk, pk_src = peaks()
    eng = a.engine or E.diffusion.DEFAULT_ENGINE
    K_prec = N
    ms_per_step = ms / a.steps
    value = world * B / (ms_per_step / 1e3)
    e2e = world * B / (ms_h / a.steps / 1e3)
    path_tflops = value / world * N * FLOP_PER_WINDOW_CALL / 1e12          # per GPU
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    W_sets = 1
    S_split = None
    if eng == "tcgen05":
        K_prec = m.precise_last_steps()
        W_sets = m.weight_sets()
        mm = re.search(r"split_steps=(\d+)", m.engine_info())
        S_split = int(mm.group(1)) if mm else K_prec
    line = {
        "metric": "motion-windows/sec (T=120, 1000-step)", "value": value, "unit": "windows/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": (f"fp16 single-pass (dithered weight sets x{W_sets}) for t>={K_prec}, "
                  + (f"fp16 activations x fp16 hi/lo weight pair for {S_split}<=t<{K_prec}, " if S_split < K_prec else "")
                  + f"fp16 hi/lo 3-term split for t<{S_split} (fp32 accumulate)"
                  if K_prec < N else "fp16 hi/lo 3-term split (fp32 accumulate)") if eng == "tcgen05" else "f32",
        "data": "synthetic", "config": cfg,      # identical in both arms; the engine's own settings are a separate key
        "engine_config": {"engine": eng, "precise_last_steps": K_prec, "split_steps": S_split, "weight_sets": W_sets},
        "e2e": {"value": e2e, "unit": "windows/s", "h2d_bytes_per_step": 2 * B * T * D * 4 * world, "d2h_bytes_per_step": B * T * D * 4 * world},
        "gpu_launches": int(launches), "clocks": clocks, "shard_check": shard_check, "engine_info": m.engine_info(),
        "path_roofline": {"bound": "tensor", "achieved": path_tflops, "peak": peak, "unit": "TFLOP/s", "frac": path_tflops / peak,
                          "scope": "whole sampling path: algorithmic 2.8507 TFLOP per 1000-step window / wall time, per GPU"},
    }
    if eng == "tcgen05":
        # every kernel of the step timed live, in isolation (20 back-to-back launches, CUDA events on the launch stream),
        # in the format of the steps that take most of the time; the DOMINANT kernel = largest launches x time share
        dom_fmt = "fp16_single" if K_prec < N else "fp16x3_split"
        half = dom_fmt == "fp16_single"
        pk_burst, hbm = pk["bf16_tflops"], pk["hbm_gbs"]
        kernels = {}
        fused_ddpm = m.launches_per_step("ddpm_update") == 0
        for name in m.KERNELS:
            cnt = m.launches_per_step(name)
            if cnt == 0:
                continue
            kms = m.time_kernel(name, B, T, half, iters=20)
            fl, by = KERNEL_FLOPS[name] * B, KERNEL_BYTES[name] * B
            if name == "out" and fused_ddpm:       # linear_out + DDPM update in one kernel: model_out never touches HBM
                by = (_ROW + KERNEL_BYTES["ddpm_update"] - 120 * 198 * 4) * B
            t_tensor, t_hbm = fl / (pk_burst * 1e12), by / (hbm * 1e9)
            bound = "tensor" if t_tensor >= t_hbm else "hbm"
            ach = fl / (kms * 1e-3) / 1e12 if bound == "tensor" else by / (kms * 1e-3) / 1e9
            kernels[name] = {"launches_per_step": cnt, "ms_per_launch": kms, "bound": bound, "achieved": ach,
                             "peak": pk_burst if bound == "tensor" else hbm, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                             "frac": ach / (pk_burst if bound == "tensor" else hbm),
                             "algorithmic_flops_per_launch": fl, "algorithmic_bytes_per_launch": by}
        # once-per-sample kernels of the path (not in the step sum): post-processing + FK on the finished windows
        try:
            ds_k = E.MotionDataStub().bind(m)
            yk = torch.rand(B, T, D, device=dev) * 2 - 1

            def _timed(fn, iters=20):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / iters

            aa_k, root_k, _h = m.postprocess(ds_k, yk, None)
            post_ms = _timed(lambda: m.postprocess(ds_k, yk, None, with_fk=True))
            fk_ms = _timed(lambda: m.fk_smpl(ds_k, root_k.reshape(-1, 3), aa_k.reshape(-1, 22, 3)))
            post_by = B * T * (198 + 66 + 3 + 3 + 66 + 88) * 4          # read x; write aa, root, head, jpos, gquat
            fk_by = B * T * (3 + 66 + 88 + 66) * 4
            line["once_per_sample_kernels"] = {
                "postprocess_fk": {"kernel": "postprocess_kernel (convert_model_res_to_data + FK)", "ms_per_launch": post_ms, "bound": "hbm",
                                   "algorithmic_bytes_per_launch": post_by, "achieved": post_by / (post_ms * 1e-3) / 1e9, "peak": hbm,
                                   "unit": "GB/s", "frac": post_by / (post_ms * 1e-3) / 1e9 / hbm},
                "fk_smpl": {"kernel": "fk_smpl_kernel", "ms_per_launch": fk_ms, "bound": "hbm", "algorithmic_bytes_per_launch": fk_by,
                            "achieved": fk_by / (fk_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": fk_by / (fk_ms * 1e-3) / 1e9 / hbm},
                "note": "one launch per sample() call (1000 diffusion steps); transcendental-heavy (6D -> matrix -> quaternion -> axis-angle per joint), "
                        "issue-bound at a few % of the HBM roofline (profiles/r2a_postprocess_kernel_full.md) and < 0.02 % of a sample() call"}
        except Exception as ex:
            line["once_per_sample_kernels"] = {"error": repr(ex)[:200]}
        tot = sum(k["launches_per_step"] * k["ms_per_launch"] for k in kernels.values())
        for k in kernels.values():
            k["share_of_step"] = k["launches_per_step"] * k["ms_per_launch"] / tot
        dom = max(kernels, key=lambda n: kernels[n]["share_of_step"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")   # dram bytes/launch from the committed ncu --set full capture
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("stage") == dom and tj.get("format") == dom_fmt and tj.get("windows") == B:
                traffic = tj.get("dram_bytes_read", 0) + tj.get("dram_bytes_write", 0)
        d = kernels[dom]
        line["roofline"] = {"bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"], "frac": d["frac"],
                            "traffic": traffic, "kernel": f"{KERNEL_NAMES[dom]} [{dom_fmt}]", "stage": dom,
                            "share_of_step": d["share_of_step"], "ms_per_launch": d["ms_per_launch"],
                            "algorithmic_flops_per_launch": d["algorithmic_flops_per_launch"],
                            "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
                            "peak_source": pk_src + ", burst figures (kernel timed alone, 20 back-to-back launches, CUDA events)"}
        line["kernels"] = kernels
        line["kernels_sum_ms_per_diffusion_step"] = tot
        # the other BASELINE metrics: denoiser forward latency (one call, split format, B=1 and B=256) ...
        line["denoiser_fwd_us"] = denoiser_latency(m, dev, T)
        line["parity_vs_reference"] = parity_vs_reference(m, dev, N)
    else:
        line["roofline"] = dict(line["path_roofline"], traffic=None, peak_source=pk_src)
    light = bool(os.environ.get("EGOEGO_BENCH_LIGHT"))      # profiler runs: headline + kernel table only
    if world == 1 and eng == "tcgen05" and not light:
        # the same workload with EVERY step in the fp32-grade 3-term split format (precise_last_steps = N): what the path costs
        # without the step-adaptive precision policy
        try:
            ms_split = all_split_ms(dev, xs, cm, B, T, N)
            line["all_split_windows_per_s"] = B / (ms_split * 1e-3)
            line["all_split_ms_per_step"] = ms_split
        except Exception as ex:
            line["all_split_windows_per_s"] = {"error": repr(ex)[:200]}
    if world == 1:                                       # reported baseline: rank 0 at N = 1 only
        line["cpu_baseline"] = cpu_baseline(32, T, N, a.cpu_seconds)
    if world == 1 and not light:
        try:
            line["next_rows"] = next_rows(dev, B, T, pk)
            line["single_window_latency"] = config1_latency(dev, T)
            line["pipeline_config4"] = pipeline_config4(dev)
        except Exception as ex:   # reported extras only; never masks the headline numbers
            line["next_rows"] = {"error": repr(ex)[:200]}
    if world == 1 and not os.environ.get("EGOEGO_BENCH_SKIP_TORCH"):
        del m
        torch.cuda.empty_cache()
        try:
            line["torch_gpu_baseline"] = torch_gpu_baseline(dev, B, T, N)
        except Exception as ex:   # reported baseline only; never masks the product numbers
            line["torch_gpu_baseline"] = {"error": repr(ex)[:200]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
