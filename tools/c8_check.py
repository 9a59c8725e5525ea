#!/usr/bin/env python
"""Cluster-of-8 multicast GEMM (EGOEGO_GEMM_C8=1) against the default CTA-pair kernel: same seed, all-fp16 steps -- the tile
arithmetic is identical (same MMA order per tile), so the samples must be bit-identical; then per-kernel and in-loop times."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import egoego_release_b200 as E
    from oracle import egoego_oracle as O
    from oracle.gen_golden import synth_x_start
    B, N = int(sys.argv[2]), int(sys.argv[3])
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                out_dim=198, timesteps=N, objective="pred_x0", max_batch=B, precise_last_steps=E.PRECISE_ALL_FP16)
    m.load_state_dict(O.init_params(0), strict=False)
    m = m.cuda()
    xs = synth_x_start(1, B, 120).cuda()
    cm = O.prep_head_condition_mask(xs.shape).cuda()
    torch.manual_seed(0)
    y = m.sample(xs, cm)
    torch.cuda.synchronize()
    print("INFO", m.engine_info(), flush=True)
    torch.save(y.cpu(), sys.argv[4])
    for name in ("qkv", "w1"):
        print("TIME", name, round(min(m.time_kernel(name, B, 120, True, iters=20) for _ in range(3)) * 1e3, 1), "us", flush=True)
    sys.exit(0)

B, N = 256, 8
outs = {}
for tag, env in (("default", {}), ("c8", {"EGOEGO_GEMM_C8": "1"})):
    e = dict(os.environ, **env)
    path = f"/tmp/c8_{tag}.pt"
    r = subprocess.run([sys.executable, __file__, "child", str(B), str(N), path], env=e, capture_output=True, text=True, timeout=120)
    print(f"[{tag}] rc={r.returncode}")
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith(("INFO", "TIME"))))
    if r.returncode != 0:
        print(r.stderr[-1500:])
        continue
    outs[tag] = torch.load(path)
if len(outs) == 2:
    print("bit-identical:", bool(torch.equal(outs["default"], outs["c8"])), "max abs diff", float((outs["default"] - outs["c8"]).abs().max()))
