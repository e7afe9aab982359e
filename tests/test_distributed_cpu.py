"""world_size-2 gloo test of the N>1 path on CPU: sharding + the final gather give the same batch as one process."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from generativemodels_b200 import distributed as D


def test_shard_bounds():
    for n in (1, 2, 5, 8, 9):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_sampler(input_noise, conditioning=None, scale=1.0):
    """Stand-in for inferer.sample: any per-sample map (each sample's trajectory is independent)."""
    out = torch.tanh(input_noise * scale)
    if conditioning is not None:
        out = out + conditioning.view(-1, 1, 1, 1)
    return out


def _worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1234)
    noise = torch.randn(total, 1, 4, 4)           # one global draw on every rank, sliced by rank
    cond = torch.arange(total, dtype=torch.float32)
    out = D.sample_sharded(_fake_sampler, noise, conditioning=cond, scale=0.5)
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


def _launch(total):
    """Two gloo ranks on a fresh local port; None if the rendezvous itself failed (e.g. the port was taken between
    probing it and binding it) so that the caller can retry on another port."""
    import queue
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, ret)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = ret.get(timeout=120)
    except queue.Empty:
        got = None
    for p in procs:
        p.join(timeout=120)
        if p.is_alive():
            p.kill()
            p.join()
    return got if all(p.exitcode == 0 for p in procs) else None


@pytest.mark.parametrize("total", [1, 2, 5])       # 1: the second rank's shard is empty
def test_sharded_sampling_matches_single_process(total):
    got = None
    for _ in range(3):
        got = _launch(total)
        if got is not None:
            break
    assert got is not None, "both ranks must finish the sharded sampling and the gather"
    torch.manual_seed(1234)
    noise = torch.randn(total, 1, 4, 4)
    want = _fake_sampler(noise, torch.arange(total, dtype=torch.float32), 0.5)
    assert torch.equal(got, want)
