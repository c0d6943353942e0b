"""linne_b200.shard -- partition a stream's blocks across ranks (one process per GPU).

Blocks are self-contained (every block re-seeds its pre-emphasis state and carries its own
coefficients), so a file -- or a corpus of files -- is sharded by
CONTIGUOUS BLOCK RANGES with no data-path collective:

  encode: rank r encodes blocks [b0_r, b1_r) of the PCM into a shard bitstream; an exclusive scan
          over the shard byte counts gives every shard's base offset behind the 30-byte header; the
          shards are then gathered (rank 0) -- the only exchange step, sizes first, bytes second.
  decode: rank 0 hops over the block size fields, ranks take contiguous block ranges of the stream;
          outputs are disjoint sample ranges, so nothing is exchanged.

Two gathers are provided:
  encode_distributed      sizes and shard bytes through `torch.distributed` (gloo in the CPU-only tests,
                          NCCL on GPUs): the portable baseline.
  encode_distributed_p2p  the B200 path: shards stay in HBM; `torch.distributed` carries only the 8-byte
                          shard sizes and the 64-byte CUDA IPC handle of rank 0's destination buffer (control
                          plane); every rank then writes its shard into that buffer at its scanned offset
                          with ONE device-to-device copy over NVLink (LINNEB200_DeviceCopy on a peer
                          mapping).  No data-path collective.
The codec object is anything with `encode(pcm, ...) -> bytes` / `decode(bytes)` (the CUDA `Product`; the
tests also drive the host simulator through the same code).
"""
from __future__ import annotations

import numpy as np

HEADER = 30


def file_ranges(num_files: int, world: int):
    """Contiguous file ranges [(lo, hi)] of a corpus over `world` ranks (per-file API calls are independent,
    so ranks take whole files; bench.py's C5 leg and the corpus tools use the same split)."""
    return [(r * num_files // world, (r + 1) * num_files // world) for r in range(world)]


def block_ranges(num_samples: int, block: int, world: int):
    """Contiguous block ranges, as sample ranges [(lo, hi)] per rank (may be empty for high ranks)."""
    nblocks = (num_samples + block - 1) // block
    per, extra = divmod(nblocks, world)
    out, b = [], 0
    for r in range(world):
        nb = per + (1 if r < extra else 0)
        lo, hi = min(b * block, num_samples), min((b + nb) * block, num_samples)
        out.append((lo, hi))
        b += nb
    return out


def patch_num_samples(header: bytes, num_samples: int) -> bytes:
    h = bytearray(header[:HEADER])
    h[14:18] = int(num_samples).to_bytes(4, "big")
    return bytes(h)


def encode_shard(codec, pcm: np.ndarray, rank: int, world: int, block: int, **fmt):
    """Blocks of this rank -> (header for the whole stream, shard bytes without a header)."""
    n = pcm.shape[1]
    lo, hi = block_ranges(n, block, world)[rank]
    if hi <= lo:
        return None, b""
    stream = codec.encode(np.ascontiguousarray(pcm[:, lo:hi]), block=block, **fmt)
    return patch_num_samples(stream[:HEADER], n), stream[HEADER:]


def exclusive_scan(sizes):
    out, acc = [], 0
    for s in sizes:
        out.append(acc)
        acc += int(s)
    return out, acc


def assemble(header: bytes, shards) -> bytes:
    return header + b"".join(shards)


def encode_distributed(codec, pcm: np.ndarray, block: int, device=None, **fmt):
    """Collective: every rank calls it with the same PCM (or at least its own range); rank 0 returns the
    whole stream, other ranks return None.  Exchange = all_gather of shard sizes + gather of bytes."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    header, shard = encode_shard(codec, pcm, rank, world, block, **fmt)
    dev = device or torch.device("cpu")
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(shard)], dtype=torch.int64, device=dev))
    sizes = [int(s.item()) for s in sizes]
    offsets, total = exclusive_scan(sizes)
    cap = max(max(sizes), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if shard:
        mine[:len(shard)] = torch.frombuffer(bytearray(shard), dtype=torch.uint8).to(dev)
    bufs = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, bufs, dst=0)
    if rank != 0:
        return None
    out = bytearray(HEADER + total)
    out[:HEADER] = header if header is not None else patch_num_samples(bytes(HEADER), pcm.shape[1])
    for r in range(world):
        out[HEADER + offsets[r]:HEADER + offsets[r] + sizes[r]] = bufs[r][:sizes[r]].cpu().numpy().tobytes()
    return bytes(out)


def stream_block_table(stream: bytes):
    """Hop over the block size fields: [(byte offset, byte size, samples)]."""
    off, table = HEADER, []
    while off + 11 <= len(stream):
        size = int.from_bytes(stream[off + 2:off + 6], "big") + 6
        table.append((off, size, int.from_bytes(stream[off + 9:off + 11], "big")))
        off += size
    return table


def decode_shard(codec, stream: bytes, rank: int, world: int):
    """Decode this rank's contiguous block range -> (first sample index, int32 [C][n_r])."""
    table = stream_block_table(stream)
    per, extra = divmod(len(table), world)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    if b1 <= b0:
        return 0, None
    first = sum(t[2] for t in table[:b0])
    count = sum(t[2] for t in table[b0:b1])
    sub = patch_num_samples(stream[:HEADER], count) + stream[table[b0][0]:table[b1 - 1][0] + table[b1 - 1][1]]
    return first, codec.decode(sub)


def encode_distributed_p2p(pcm: np.ndarray, block: int, bits=16, rate=44100, preset=0, device=None, to_host=True):
    """Collective: every rank encodes its contiguous block range ON ITS GPU and puts the shard into rank 0's
    device buffer over NVLink.  Rank 0 returns (DeviceBuffer holding the whole stream, its size[, bytes]);
    other ranks return None.  Needs the NCCL (or gloo) process group for the control messages only."""
    import torch
    import torch.distributed as dist
    from .api import EncoderSession, DeviceBuffer, PeerMapping
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = device or torch.device("cuda", torch.cuda.current_device())
    nch, n = pcm.shape
    lo, hi = block_ranges(n, block, world)[rank]
    count = hi - lo
    shard_size, d_shard = 0, None
    if count > 0:
        stride = (count + 4 + 3) // 4 * 4
        d_pcm = torch.zeros((nch, stride), dtype=torch.int32, device=dev)
        d_pcm[:, :count].copy_(torch.from_numpy(np.ascontiguousarray(pcm[:, lo:hi])))
        cap = HEADER + 2 * nch * count * 4 + 65536
        d_shard = torch.zeros(cap + 64, dtype=torch.uint8, device=dev)
        enc = EncoderSession(nch, bits=bits, rate=rate, block=block, preset=preset)
        try:
            torch.cuda.synchronize(dev)
            shard_size = enc.encode_whole_resident(d_pcm.data_ptr(), stride, count, d_shard.data_ptr(), cap) - HEADER
        finally:
            enc.close()
    # control plane: shard sizes (exclusive scan on every rank) and the IPC handle of the destination
    ctl_dev = dev if dist.get_backend() == "nccl" else torch.device("cpu")
    sizes_t = [torch.zeros(1, dtype=torch.int64, device=ctl_dev) for _ in range(world)]
    dist.all_gather(sizes_t, torch.tensor([shard_size], dtype=torch.int64, device=ctl_dev))
    sizes = [int(t.item()) for t in sizes_t]
    offsets, total = exclusive_scan(sizes)
    handle_t = torch.zeros(64, dtype=torch.uint8, device=ctl_dev)
    dest = None
    if rank == 0:
        dest = DeviceBuffer(HEADER + total + 64)
        handle_t.copy_(torch.frombuffer(bytearray(dest.ipc_handle()), dtype=torch.uint8))
    dist.broadcast(handle_t, src=0)
    # data plane: one device-to-device copy per rank, straight into rank 0's HBM
    if rank == 0:
        if shard_size:
            dest.lib.LINNEB200_DeviceCopy(dest.ptr + HEADER + offsets[0], d_shard.data_ptr() + HEADER, shard_size)
            hdr = bytearray(d_shard[:HEADER].cpu().numpy().tobytes())
        else:
            hdr = bytearray(HEADER)
        dest.upload(patch_num_samples(bytes(hdr), n), 0)
    elif shard_size:
        peer = PeerMapping(bytes(handle_t.cpu().numpy().tobytes()))
        try:
            peer.put(HEADER + offsets[rank], d_shard.data_ptr() + HEADER, shard_size)
        finally:
            peer.close()
    dist.barrier()
    if rank != 0:
        return None
    if to_host:
        return dest, HEADER + total, dest.download(HEADER + total)
    return dest, HEADER + total
