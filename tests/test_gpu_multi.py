"""Multi-GPU sharding on real devices: one process per GPU, contiguous block ranges,
exclusive scan of shard byte counts, shards written into rank 0's HBM over NVLink through a CUDA IPC peer
mapping (no data-path collective), decode by block range with no exchange.  Needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, pcm, preset, block, results):
    import torch
    import torch.distributed as dist
    from linne_b200 import Product, shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = shard.encode_distributed_p2p(pcm, block, preset=preset)
        stream = None
        if rank == 0:
            dest, size, stream = out
            results["stream"] = stream
            dest.free()
        # decode: every rank takes its own contiguous block range of the gathered stream, nothing is exchanged
        box = [stream]
        dist.broadcast_object_list(box, src=0)
        first, got = shard.decode_shard(Product(), box[0], rank, world)
        results[f"dec{rank}"] = (first, got)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("preset", (0, 7))
def test_p2p_sharded_encode_and_range_decode(preset):
    import torch.multiprocessing as mp
    from linne_b200 import Product
    world, block = 2, 4096
    pcm = harness.synth_pcm(n=block * 9 + 1500, channels=2, bits=16, seed=41)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, pcm, preset, block, results), nprocs=world, join=True)
    stream = results["stream"]
    whole = Product().encode(pcm, preset=preset, block=block)
    assert stream == whole                                   # shards concatenate to the single-call stream
    assert np.array_equal(harness.Oracle().decode(stream), pcm)
    out = np.zeros_like(pcm)
    for r in range(world):
        first, got = results[f"dec{r}"]
        out[:, first:first + got.shape[1]] = got
    assert np.array_equal(out, pcm)
