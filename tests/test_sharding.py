"""Multi-rank sharding by contiguous block ranges: host logic on the CPU with the
gloo backend (world_size 2), plus the single-process equivalences that make sharding legal."""
import os
import socket

import numpy as np
import pytest

import harness
from linne_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


# ---- a corpus sharded by contiguous FILE ranges, several files per call (LINNEB200_EncodeFilesResident) ----
def test_file_ranges_cover_the_corpus():
    for files, world in ((1000, 8), (5, 2), (3, 4), (1, 2)):
        r = shard.file_ranges(files, world)
        assert r[0][0] == 0 and r[-1][1] == files and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def _encode_files_batch(codec, files, block, preset):
    """one LINNEB200_EncodeFilesResident call over `files` (int32 [C][n] each) on the simulator -> [stream bytes]"""
    import ctypes as C
    from harness import LINNEEncodeParameter, LINNEEncoderConfig, OK
    from linne_b200.api import FileDesc
    if not files:
        return []
    L = codec.lib
    nch = files[0].shape[0]
    starts, off = [], 0
    for f in files:
        starts.append(off); off += (f.shape[1] + 3) // 4 * 4
    planes = np.zeros((nch, max(off, 4)), np.int32)
    for f, s in zip(files, starts):
        planes[:, s:s + f.shape[1]] = f
    cap = sum(30 + f.size * 4 + 4096 for f in files)
    out = np.zeros(cap + 64, np.uint8)
    desc = (FileDesc * len(files))(*[FileDesc(s, f.shape[1], 0, 0, 0) for f, s in zip(files, starts)])
    enc = L.LINNEEncoder_Create(C.byref(LINNEEncoderConfig(nch, block, 3, 128)), None, 0)
    try:
        assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(LINNEEncodeParameter(nch, 16, 44100, block, preset, 1, 0, 0))) == OK
        total = C.c_uint32(0)
        assert L.LINNEB200_EncodeFilesResident(enc, C.c_void_p(planes.ctypes.data), planes.shape[1], desc, len(files),
                                               C.c_void_p(out.ctypes.data), cap, C.byref(total)) == OK
        return [out[d.out_offset:d.out_offset + d.out_size].tobytes() for d in desc]
    finally:
        L.LINNEEncoder_Destroy(enc)


def _corpus_worker(rank, world, port, corpus, results):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        codec = harness.HostSim()
        lo, hi = shard.file_ranges(len(corpus), world)[rank]
        mine = _encode_files_batch(codec, corpus[lo:hi], 2048, 3)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)              # control plane only: the streams stay where they were made
        if rank == 0:
            results["streams"] = [s for part in gathered for s in part]
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_corpus_by_file_ranges(hostsim, oracle):
    import torch.multiprocessing as mp
    corpus = [harness.synth_pcm(n=n, channels=2, bits=16, seed=60 + i) for i, n in enumerate((2048 * 2 + 100, 2048, 3000, 2048 * 3, 700))]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_corpus_worker, args=(2, port, corpus, results), nprocs=2, join=True)
    streams = results["streams"]
    assert len(streams) == len(corpus)
    for f, st in zip(corpus, streams):
        assert st == hostsim.encode(f, preset=3, block=2048)
        assert np.array_equal(oracle.decode(st), f)


def test_turnstile_orders_the_ranges_of_a_device():
    """LnbTurnstile (csrc/host/lnb_host_util.c): range k may pass a phase only after range k - stride has; passing is
    idempotent and a range that bails out early unblocks its successors.  Driven from Python threads against the
    host-simulator build of the same host C code (no GPU involved)."""
    import ctypes as C
    import threading
    import time
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "liblinne_hostsim.so"))
    buf = C.create_string_buffer(512)                  # room for the struct (mutex + condition variable + flags)
    for f in (lib.lnb_turnstile_init, lib.lnb_turnstile_wait, lib.lnb_turnstile_pass, lib.lnb_turnstile_destroy):
        f.restype = None
    for stride, ranges in ((1, 6), (2, 8)):
        lib.lnb_turnstile_init(buf, C.c_uint32(stride))
        order = {0: [], 1: []}
        lock = threading.Lock()

        def work(k):
            for phase in (0, 1):
                lib.lnb_turnstile_wait(buf, C.c_int(phase), C.c_uint32(k))
                with lock:
                    order[phase].append(k)
                time.sleep(0.002 * ((k * 7 + phase) % 3))
                if k == 3 and phase == 0:              # range 3 gives up after its upload: both phases are passed for it
                    lib.lnb_turnstile_pass(buf, C.c_int(0), C.c_uint32(k))
                    lib.lnb_turnstile_pass(buf, C.c_int(1), C.c_uint32(k))
                    return
                lib.lnb_turnstile_pass(buf, C.c_int(phase), C.c_uint32(k))
                lib.lnb_turnstile_pass(buf, C.c_int(phase), C.c_uint32(k))      # idempotent
        threads = [threading.Thread(target=work, args=(k,)) for k in reversed(range(ranges))]
        for t in threads: t.start()
        for t in threads: t.join(timeout=20)
        assert not any(t.is_alive() for t in threads), "turnstile deadlock"
        lib.lnb_turnstile_destroy(buf)
        for phase in (0, 1):
            seen = order[phase]
            assert sorted(seen) == [k for k in range(ranges) if not (phase == 1 and k == 3)]
            for lane in range(stride):                 # ranges that share a device went through in range order
                mine = [k for k in seen if k % stride == lane]
                assert mine == sorted(mine), (stride, phase, seen)


def test_plan_ranges_pipeline_depth(monkeypatch):
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "liblinne_hostsim.so"))
    lib.lnb_plan_ranges.restype = C.c_uint32
    monkeypatch.delenv("LINNE_B200_PIPELINE", raising=False)
    assert lib.lnb_plan_ranges(C.c_uint32(44), C.c_uint32(1)) == 1           # a clip: the handle's own device, no ranges
    assert lib.lnb_plan_ranges(C.c_uint32(3072), C.c_uint32(1)) == 2
    assert lib.lnb_plan_ranges(C.c_uint32(15504), C.c_uint32(1)) == 4        # the 1-hour stream
    assert lib.lnb_plan_ranges(C.c_uint32(15504), C.c_uint32(8)) == 8
    assert lib.lnb_plan_ranges(C.c_uint32(5), C.c_uint32(8)) == 2            # at least two blocks per range
    monkeypatch.setenv("LINNE_B200_PIPELINE", "3")
    assert lib.lnb_plan_ranges(C.c_uint32(15504), C.c_uint32(2)) == 6


class _Ranges:
    def __init__(self, n):
        self.n = n

    def __enter__(self):
        self.old = os.environ.get("LINNE_B200_GPUS")
        os.environ["LINNE_B200_GPUS"] = str(self.n)

    def __exit__(self, *a):
        if self.old is None:
            del os.environ["LINNE_B200_GPUS"]
        else:
            os.environ["LINNE_B200_GPUS"] = self.old


@pytest.mark.parametrize("ranges", [2, 5])
def test_in_library_block_ranges_on_the_host_simulator(hostsim, oracle, ranges):
    """EncodeWhole / DecodeWhole over several block ranges behind one handle (child handles, one host thread per range,
    host-side scan of the shard sizes, transfers taking turns): the host C code of the product, run against the CPU
    stand-in of the shim.  Same bytes and samples as the single-range call; a damaged stream reports its first error."""
    pcm = harness.synth_pcm(n=1024 * 13 + 300, channels=2, bits=16, seed=91)
    single = hostsim.encode(pcm, preset=1, block=1024)
    with _Ranges(ranges):
        sharded = hostsim.encode(pcm, preset=1, block=1024)
        back = hostsim.decode(single)
    assert sharded == single
    assert np.array_equal(back, pcm)
    assert np.array_equal(oracle.decode(sharded), pcm)
    sizes, off = [], 30
    while off < len(single):
        size = int.from_bytes(single[off + 2:off + 6], "big") + 6
        sizes.append(size); off += size
    bad = bytearray(single)
    bad[30 + sum(sizes[:4]) + 30] ^= 0x10                       # block 4: CRC mismatch
    bad[30 + sum(sizes[:9]) + 30] ^= 0x10                       # block 9 too: the earlier one decides
    rc1, out1 = hostsim.decode(bytes(bad), return_code=True, fill=-7)
    with _Ranges(ranges):
        rc2, out2 = hostsim.decode(bytes(bad), return_code=True, fill=-7)
    assert rc1 == rc2 == harness.DATA_CORRUPTION
    assert np.array_equal(out1[:, :4 * 1024], pcm[:, :4 * 1024]) and np.array_equal(out2[:, :4 * 1024], pcm[:, :4 * 1024])
