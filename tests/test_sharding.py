"""Multi-rank sharding by contiguous block ranges: host logic on the CPU with the
gloo backend (world_size 2), plus the single-process equivalences that make sharding legal."""
import os
import socket

import numpy as np
import pytest

import harness
from linne_b200 import shard


def test_block_ranges_cover_everything():
    for n, block, world in ((441000, 10240, 8), (5000, 2048, 2), (1000, 4096, 4), (10240 * 3, 10240, 2)):
        r = shard.block_ranges(n, block, world)
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(lo % block == 0 or lo == n for lo, _ in r)


@pytest.mark.parametrize("world", (2, 3))
def test_shards_concatenate_to_the_single_call_stream(hostsim, world):
    pcm = harness.synth_pcm(n=2048 * 5 + 700, channels=2, bits=16, seed=31)
    whole = hostsim.encode(pcm, preset=2, block=2048)
    parts = [shard.encode_shard(hostsim, pcm, r, world, 2048, preset=2) for r in range(world)]
    assert shard.assemble(parts[0][0], [p[1] for p in parts]) == whole
    # decode by block ranges: disjoint sample ranges, no exchange
    out = np.zeros_like(pcm)
    for r in range(world):
        first, got = shard.decode_shard(hostsim, whole, r, world)
        if got is not None:
            out[:, first:first + got.shape[1]] = got
    assert np.array_equal(out, pcm)


def _worker(rank, world, port, pcm, results):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        codec = harness.HostSim()
        stream = shard.encode_distributed(codec, pcm, 2048, preset=0)
        if rank == 0:
            results["stream"] = stream
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_encode(hostsim, oracle):
    import torch.multiprocessing as mp
    pcm = harness.synth_pcm(n=2048 * 4 + 300, channels=2, bits=16, seed=33)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, pcm, results), nprocs=2, join=True)
    stream = results["stream"]
    assert stream == hostsim.encode(pcm, preset=0, block=2048)
    assert np.array_equal(oracle.decode(stream), pcm)
