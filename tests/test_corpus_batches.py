"""Corpus batches ("concatenate files' block lists, keeping per-file boundaries"):
LINNEB200_EncodeFilesResident / LINNEB200_DecodeFilesResident run the blocks of several files through the kernels as
one batch.  Every stream must be byte-identical with what EncodeWhole writes for that file alone (and so with the
oracle's), every file must decode to its PCM, and a damaged file must not disturb its neighbours.

CPU: the host code on the simulator (its "device" memory is host memory).  GPU: the product, buffers in HBM."""
import ctypes as C

import numpy as np
import pytest

import harness
from harness import OK, DATA_CORRUPTION, INSUFFICIENT_BUFFER, INVALID_ARGUMENT, LINNEEncodeParameter, LINNEEncoderConfig, LINNEDecoderConfig
from linne_b200.api import FileDesc

BLOCK = 2048
LENGTHS = [BLOCK * 3 + 700, BLOCK * 2, 900, BLOCK * 5 + 1, BLOCK]      # tail blocks, whole blocks, a file shorter than a block


def make_corpus(channels, bits, seed):
    files = [harness.synth_pcm(n=n, channels=channels, bits=bits, seed=seed + i) for i, n in enumerate(LENGTHS)]
    starts, off = [], 16                                    # files need not start at sample 0 nor touch each other
    for f in files:
        starts.append(off); off += f.shape[1] + 24
    off = (off + 3) // 4 * 4
    planes = np.zeros((channels, off), np.int32)
    for f, s in zip(files, starts):
        planes[:, s:s + f.shape[1]] = f
    return files, starts, planes


class HostMem:
    """`device` memory of the simulator = host memory"""
    def __init__(self, arr): self.arr = np.ascontiguousarray(arr)
    @property
    def ptr(self): return self.arr.ctypes.data
    def get(self): return self.arr
    def put(self, a): self.arr[...] = a


class CudaMem:
    def __init__(self, arr):
        import torch
        self.t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    @property
    def ptr(self): return self.t.data_ptr()
    def get(self): return self.t.cpu().numpy()
    def put(self, a):
        import torch
        self.t.copy_(torch.from_numpy(np.ascontiguousarray(a)))


def run_batch(codec, Mem, channels, bits, preset, oracle):
    L = codec.lib
    files, starts, planes = make_corpus(channels, bits, seed=300 + preset)
    stride = planes.shape[1]
    d_pcm = Mem(planes)
    cap = sum(30 + f.size * 4 + 4096 for f in files)
    d_out = Mem(np.zeros(cap + 64, np.uint8))
    desc = (FileDesc * len(files))(*[FileDesc(s, f.shape[1], 0, 0, 0) for f, s in zip(files, starts)])
    enc = L.LINNEEncoder_Create(C.byref(LINNEEncoderConfig(channels, BLOCK, 3, 128)), None, 0)
    dec = L.LINNEDecoder_Create(C.byref(LINNEDecoderConfig(channels, 3, 128, 1)), None, 0)
    try:
        ms = 1 if channels >= 2 else 0
        assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(LINNEEncodeParameter(channels, bits, 44100, BLOCK, preset, ms, 0, 0))) == OK
        total = C.c_uint32(0)
        assert L.LINNEB200_EncodeFilesResident(enc, C.c_void_p(d_pcm.ptr), stride, desc, len(files), C.c_void_p(d_out.ptr), cap,
                                               C.byref(total)) == OK
        image = d_out.get()
        pos = 0
        for f, d in zip(files, desc):
            assert d.out_offset == pos and d.out_size > 30
            got = image[d.out_offset:d.out_offset + d.out_size].tobytes()
            assert got == codec.encode(f, bits=bits, preset=preset, block=BLOCK), "differs from the single-file call"
            assert np.array_equal(oracle.decode(got), f)
            pos += d.out_size
        assert total.value == pos
        # too small an output buffer
        small = C.c_uint32(0)
        assert L.LINNEB200_EncodeFilesResident(enc, C.c_void_p(d_pcm.ptr), stride, desc, len(files), C.c_void_p(d_out.ptr),
                                               desc[0].out_size + desc[1].out_size + 10, C.byref(small)) == INSUFFICIENT_BUFFER
        assert L.LINNEB200_EncodeFilesResident(enc, C.c_void_p(d_pcm.ptr), stride, desc, 0, C.c_void_p(d_out.ptr), cap,
                                               C.byref(small)) == INVALID_ARGUMENT
        assert L.LINNEB200_EncodeFilesResident(enc, C.c_void_p(d_pcm.ptr), stride, desc, len(files), C.c_void_p(d_out.ptr), cap,
                                               C.byref(total)) == OK

        # decode the batch into fresh planes: files land where the table says, nothing else is touched
        d_back = Mem(np.full_like(planes, -7))
        assert L.LINNEB200_DecodeFilesResident(dec, C.c_void_p(d_out.ptr), total.value, desc, len(files), C.c_void_p(d_back.ptr), stride) == OK
        back = d_back.get()
        want = np.full_like(planes, -7)
        for f, s in zip(files, starts):
            want[:, s:s + f.shape[1]] = f
        assert np.array_equal(back, want)
        assert all(d.status == OK for d in desc)

        # one damaged file: its status says so, the call returns it, the other files still decode
        bad = image.copy()
        bad[desc[2].out_offset + 60] ^= 0x20
        d_out.put(bad)
        d_back.put(np.full_like(planes, -7))
        assert L.LINNEB200_DecodeFilesResident(dec, C.c_void_p(d_out.ptr), total.value, desc, len(files), C.c_void_p(d_back.ptr), stride) == DATA_CORRUPTION
        assert [d.status for d in desc] == [OK, OK, DATA_CORRUPTION, OK, OK]
        back = d_back.get()
        for i, (f, s) in enumerate(zip(files, starts)):
            if i != 2:
                assert np.array_equal(back[:, s:s + f.shape[1]], f)
        # a file whose PCM range is too small for its stream
        d_out.put(image)
        desc[1].num_samples -= 1
        assert L.LINNEB200_DecodeFilesResident(dec, C.c_void_p(d_out.ptr), total.value, desc, len(files), C.c_void_p(d_back.ptr), stride) == INSUFFICIENT_BUFFER
        assert desc[1].status == INSUFFICIENT_BUFFER and desc[0].status == OK
    finally:
        L.LINNEEncoder_Destroy(enc)
        L.LINNEDecoder_Destroy(dec)


def run_packed_batch(codec, channels, bits, preset, oracle):
    """host buffers: the files' WAV data chunks back to back in, the streams back to back out, and back again"""
    L = codec.lib
    u8p = C.POINTER(C.c_uint8)
    files, _, _ = make_corpus(channels, bits, seed=500 + preset)
    packed = b"".join(harness.pack_pcm(f, bits) for f in files)
    src = np.frombuffer(packed, np.uint8).copy()
    cap = sum(30 + f.size * 4 + 4096 for f in files)
    out = np.zeros(cap + 64, np.uint8)
    desc = (FileDesc * len(files))(*[FileDesc(0, f.shape[1], 0, 0, 0) for f in files])
    enc = L.LINNEEncoder_Create(C.byref(LINNEEncoderConfig(channels, BLOCK, 3, 128)), None, 0)
    dec = L.LINNEDecoder_Create(C.byref(LINNEDecoderConfig(channels, 3, 128, 1)), None, 0)
    try:
        ms = 1 if channels >= 2 else 0
        assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(LINNEEncodeParameter(channels, bits, 44100, BLOCK, preset, ms, 0, 0))) == OK
        total = C.c_uint32(0)
        assert L.LINNEB200_EncodeFilesPacked(enc, src.ctypes.data_as(u8p), desc, len(files), out.ctypes.data_as(u8p), cap, C.byref(total)) == OK
        frames = 0
        for f, d in zip(files, desc):
            assert d.first_sample == frames
            assert out[d.out_offset:d.out_offset + d.out_size].tobytes() == codec.encode(f, bits=bits, preset=preset, block=BLOCK)
            frames += f.shape[1]
        assert total.value == desc[len(files) - 1].out_offset + desc[len(files) - 1].out_size
        back = np.zeros(len(packed), np.uint8)
        assert L.LINNEB200_DecodeFilesPacked(dec, out.ctypes.data_as(u8p), total.value, desc, len(files), back.ctypes.data_as(u8p)) == OK
        assert back.tobytes() == packed
        assert L.LINNEB200_EncodeFilesPacked(enc, src.ctypes.data_as(u8p), desc, len(files), out.ctypes.data_as(u8p), 1000, C.byref(total)) == INSUFFICIENT_BUFFER
    finally:
        L.LINNEEncoder_Destroy(enc)
        L.LINNEDecoder_Destroy(dec)


@pytest.mark.parametrize("channels,bits,preset", [(2, 16, 1), (1, 24, 6), (2, 8, 2)])
def test_corpus_batches_host_buffers_hostsim(hostsim, oracle, channels, bits, preset):
    run_packed_batch(hostsim, channels, bits, preset, oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("channels,bits,preset", [(2, 16, 1), (1, 24, 6), (8, 24, 4)])
def test_corpus_batches_host_buffers_gpu(gpu, oracle, channels, bits, preset):
    run_packed_batch(gpu, channels, bits, preset, oracle)


@pytest.mark.parametrize("channels,bits,preset", [(2, 16, 0), (1, 24, 5), (2, 16, 7)])
def test_corpus_batches_hostsim(hostsim, oracle, channels, bits, preset):
    run_batch(hostsim, HostMem, channels, bits, preset, oracle)


def test_corpus_batches_in_several_chunks(hostsim, oracle, monkeypatch):
    """a scratch budget of 1 MiB cuts the batch into chunks of a few blocks: files straddle chunk boundaries"""
    monkeypatch.setenv("LINNE_B200_SCRATCH_MB", "1")
    run_batch(hostsim, HostMem, 2, 16, 4, oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("channels,bits,preset", [(2, 16, 0), (1, 24, 5), (2, 16, 7), (8, 24, 3)])
def test_corpus_batches_gpu(gpu, oracle, channels, bits, preset):
    run_batch(gpu, CudaMem, channels, bits, preset, oracle)


@pytest.mark.gpu
def test_corpus_batch_takes_the_throughput_decoder(gpu, oracle):
    """a batch of many files is one large decode: above the switch-over point it runs on the throughput kernels"""
    import torch
    L = gpu.lib
    pcm = harness.synth_pcm(n=2048 * 20, channels=2, bits=16, seed=77)
    nfiles, n = 140, pcm.shape[1]                                 # 140 files x 20 blocks of 2048 = 2800 blocks
    planes = np.ascontiguousarray(np.tile(pcm, (1, nfiles)))
    d_pcm = torch.from_numpy(planes).cuda()
    cap = nfiles * (30 + pcm.size * 4 + 4096)
    d_out = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
    desc = (FileDesc * nfiles)(*[FileDesc(i * n, n, 0, 0, 0) for i in range(nfiles)])
    enc = L.LINNEEncoder_Create(C.byref(LINNEEncoderConfig(2, 2048, 3, 128)), None, 0)
    from linne_b200 import DecoderSession
    dec = DecoderSession(channels=2)
    try:
        assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(LINNEEncodeParameter(2, 16, 44100, 2048, 6, 1, 0, 0))) == OK
        total = C.c_uint32(0)
        assert L.LINNEB200_EncodeFilesResident(enc, C.c_void_p(d_pcm.data_ptr()), planes.shape[1], desc, nfiles,
                                               C.c_void_p(d_out.data_ptr()), cap, C.byref(total)) == OK
        one = gpu.encode(pcm, preset=6, block=2048)
        image = d_out[:total.value].cpu().numpy()
        assert all(image[d.out_offset:d.out_offset + d.out_size].tobytes() == one for d in desc)
        d_back = torch.zeros_like(d_pcm)
        dec.set_profiling(True)
        assert L.LINNEB200_DecodeFilesResident(dec.h, C.c_void_p(d_out.data_ptr()), total.value, desc, nfiles,
                                               C.c_void_p(d_back.data_ptr()), planes.shape[1]) == OK
        assert {"tp_entropy", "tp_synth"} <= set(dec.stage_stats())
        assert torch.equal(d_back, d_pcm)
    finally:
        L.LINNEEncoder_Destroy(enc)
        dec.close()
