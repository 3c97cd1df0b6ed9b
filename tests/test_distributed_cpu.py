"""CPU, world_size 2 over gloo: the only collectives on the path — the start-up weight broadcast and the end-of-run
counter gather — plus the prompt partition every rank derives locally."""
from __future__ import annotations

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusion_spacetime_attn_b200.pipeline import broadcast_weights, shard_prompts

    torch.manual_seed(100 + rank)  # ranks start from DIFFERENT weights; after the broadcast they must equal rank 0's
    net = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.GroupNorm(4, 64), torch.nn.Linear(64, 8))
    net.register_buffer("sched", torch.randn(10))
    sent = broadcast_weights(net, src=0, bucket_bytes=8 << 10)  # small buckets: several collectives
    flat = torch.cat([t.reshape(-1) for t in list(net.parameters()) + list(net.buffers())])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    mine = shard_prompts(11, rank, world)
    counts = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(mine)]))
    q.put((rank, sent, bool(all(torch.equal(g, gathered[0]) for g in gathered)), mine, [int(c) for c in counts]))
    dist.destroy_process_group()


def test_weight_broadcast_and_prompt_partition_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[2] for r in res), "weights differ after the broadcast"
    assert res[0][1] == res[1][1] > 0
    assert res[0][3] == [0, 2, 4, 6, 8, 10] and res[1][3] == [1, 3, 5, 7, 9]
    assert res[0][4] == [6, 5]
