"""Host-side logic of the N>1 path on CPU: world_size-2 gloo processes shard a batch by global window id and
all-gather the finished windows; the result must equal the single-process run (sampler replaced by a deterministic
function of the global window id -- the CUDA engine itself is covered by the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egoego_release_b200.parallel import average_flat_gradients, sample_sharded, shard_range


def _fake_sample(xs, cm, offset):
    ids = torch.arange(offset, offset + xs.shape[0], dtype=torch.float32).view(-1, 1, 1)
    return xs * (1 - cm) + cm * torch.sin(ids + torch.arange(xs.shape[2]).float().view(1, 1, -1))


def _fake_post(windows, offset):
    """Stand-in for smpl_post_fn: a per-window reduction to 69 channels that also depends on the global window id."""
    ids = torch.arange(offset, offset + windows.shape[0], dtype=torch.float32).view(-1, 1, 1)
    return windows[..., :69] * 2.0 + ids


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    xs = torch.randn(B, 6, 198, generator=g)
    cm = (torch.rand(B, 6, 198, generator=g) > 0.5).float()
    out = sample_sharded(_fake_sample, xs, cm)
    packed = sample_sharded(_fake_sample, xs, cm, post_fn=_fake_post)
    if rank == 0:
        q.put((out, packed))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("B", [8, 7])
def test_sharded_sampling_equals_single_process(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, packed = q.get(timeout=300)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    xs = torch.randn(B, 6, 198, generator=g)
    cm = (torch.rand(B, 6, 198, generator=g) > 0.5).float()
    assert torch.equal(out, _fake_sample(xs, cm, 0))
    assert packed.shape == (B, 6, 69) and torch.equal(packed, _fake_post(_fake_sample(xs, cm, 0), 0))    # gather of post-processed params


def test_shard_range_partitions():
    for B in (1, 7, 8, 256, 2048):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == B
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)          # what the CUDA backward would leave in the flat buffer
    out = average_flat_gradients(flat)
    assert out is flat
    if rank == 0:
        q.put(flat.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_average_over_ranks():
    """set_grad_sync's reduction (one all-reduce of the flat gradient buffer, mean over ranks) with two gloo processes; outside a
    process group it leaves the buffer untouched."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert torch.equal(got, torch.arange(10, dtype=torch.float32) * 1.5)
    alone = torch.ones(4)
    assert torch.equal(average_flat_gradients(alone), torch.ones(4))


def test_grad_sync_switches():
    """set_grad_sync / no_grad_sync on the mirror (pure host state; the reduction itself is average_flat_gradients)."""
    import egoego_release_b200 as E
    m = E.CondGaussianDiffusion(d_feats=198, d_model=512, n_dec_layers=4, n_head=4, d_k=256, d_v=256, max_timesteps=121,
                                out_dim=198, timesteps=10, objective="pred_x0")
    assert getattr(m, "_grad_sync", None) is None
    m.set_grad_sync()
    assert m._grad_sync == (None, True)
    with m.no_grad_sync():
        assert m._grad_sync is None
    assert m._grad_sync == (None, True)
    m.set_grad_sync(enabled=False)
    assert m._grad_sync is None
