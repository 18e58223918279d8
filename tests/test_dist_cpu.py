"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: one broadcast of the packed weight blob from rank 0,
contiguous batch sharding, final gather.  The GPU path differs only in the backend (nccl) and tensor device."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mug_diffusion_b200 import dist as mdist
from mug_diffusion_b200 import synth
from mug_diffusion_b200.config import DecoderConfig, ModelConfig, UNetConfig
from mug_diffusion_b200.packer import pack_model

SMALL = ModelConfig(unet=UNetConfig(model_channels=32, channel_mult=(1, 2), num_res_blocks=1, attention_resolutions=(2,),
                                    audio_channels=(32, 32), context_dim=32, num_heads=2),
                    decoder=DecoderConfig(middle_channels=32, channel_mult=(1, 2), num_res_blocks=1, num_groups=8))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd = synth.synthetic_state_dict(64, SMALL) if rank == 0 else None
        blob = mdist.broadcast_blob(sd, SMALL, torch.device("cpu"))
        ref = pack_model(synth.synthetic_state_dict(64, SMALL), SMALL.unet, SMALL.decoder)
        same = torch.equal(blob.data, ref.data) and blob.entries.keys() == ref.entries.keys() and blob.meta == ref.meta
        off_ok = all(blob.offset(k) == ref.offset(k) for k in ref.entries)
        # batch sharding: 5 samples over 2 ranks -> 3 + 2, contiguous, covering everything once
        x = torch.arange(5 * 3, dtype=torch.float32).view(5, 3)
        (mine,) = mdist.shard_batch([x], rank, world)
        sizes = [mdist.shard_range(5, r, world)[1] - mdist.shard_range(5, r, world)[0] for r in range(world)]
        back = mdist.gather_batch(mine * 2, sizes, dst=0)
        gather_ok = True if rank != 0 else torch.equal(back, x * 2)
        q.put((rank, same, off_ok, tuple(mine.shape), gather_ok))
    finally:
        dist.destroy_process_group()


def test_broadcast_blob_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] and r[2] and r[4] for r in res)
    assert res[0][3] == (3, 3) and res[1][3] == (2, 3)


def test_shard_range_partitions_exactly():
    for total in (1, 4, 7, 256):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker_sharded(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, L = 5, 32
        shapes = dict(x_T=(B, 16, L), c=(B, 128, 21), uc=(B, 128, 21), w0=(B, 256, L), w1=(B, 512, L // 2), w2=(B, 512, L // 4), w3=(B, 512, L // 8))
        req = None
        if rank == 0:
            gen = torch.Generator().manual_seed(3)
            req = dict(x_T=torch.randn(shapes["x_T"], generator=gen), c=torch.randn(shapes["c"], generator=gen), uc=torch.randn(shapes["uc"], generator=gen),
                       w=[torch.randn(shapes[f"w{i}"], generator=gen) for i in range(4)])

        def fake_sampler(x, c, uc, w):          # any per-sample function of all inputs: proves every shard got the right rows
            return x.sum((1, 2), keepdim=True) + c.mean((1, 2), keepdim=True) - uc.amax((1, 2), keepdim=True) + sum(t.sum((1, 2), keepdim=True) for t in w)

        got = mdist.sample_sharded(fake_sampler, req, shapes, torch.device("cpu"))
        ok = True
        if rank == 0:
            ok = torch.allclose(got, fake_sampler(req["x_T"], req["c"], req["uc"], req["w"]), rtol=1e-6, atol=1e-5) and got.shape[0] == B
        else:
            ok = got is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sample_sharded_scatter_run_gather_world2():
    """the public multi-rank request call: rank 0 holds the request, shards are scattered (uneven: 3 + 2), results come back in batch order"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
