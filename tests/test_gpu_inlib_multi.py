"""Several GPUs behind ONE handle, inside the C library (include/linne_b200.h: LINNEB200_*SetDevices,
LINNE_B200_GPUS): EncodeWhole / DecodeWhole cut the stream's blocks into contiguous ranges, one child handle, device and
host thread per range; shard sizes are scanned on the host and every shard lands at its place in the caller's buffer.

Runs on whatever GPU count is visible: with more ranges than devices the ranges share devices round-robin, so a
one-GPU box walks the same code (threads, child handles, scan, placement) as an eight-GPU one.  The bar is the
single-device result, byte for byte."""
import os

import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu


class _Gpus:
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


@pytest.mark.parametrize("ranges", [2, 3, 8])
@pytest.mark.parametrize("preset", [0, 7])
def test_sharded_encode_is_byte_identical_and_decodes(gpu, oracle, ranges, preset):
    pcm = harness.synth_pcm(n=4096 * 19 + 1500, channels=2, bits=16, seed=23 + preset)       # 19 full blocks + a tail
    single = gpu.encode(pcm, preset=preset, block=4096)
    with _Gpus(ranges):
        sharded = gpu.encode(pcm, preset=preset, block=4096)
        back = gpu.decode(single)
    assert sharded == single
    assert np.array_equal(back, pcm)
    assert np.array_equal(oracle.decode(sharded), pcm)


def test_sharded_formats(gpu, oracle):
    pcm = harness.synth_pcm(n=2048 * 11 + 700, channels=8, bits=24, seed=77)
    want = oracle.encode(pcm, bits=24, preset=5, block=2048)
    single = gpu.encode(pcm, bits=24, preset=5, block=2048)
    with _Gpus(4):
        assert gpu.encode(pcm, bits=24, preset=5, block=2048) == single
        assert np.array_equal(gpu.decode(want), pcm)


def test_sharded_decode_reports_the_first_error_in_stream_order(gpu, oracle):
    pcm = harness.synth_pcm(n=2048 * 16, channels=2, bits=16, seed=31)
    good = oracle.encode(pcm, preset=0, block=2048)
    sizes, off = [], 30
    while off < len(good):
        size = int.from_bytes(good[off + 2:off + 6], "big") + 6
        sizes.append(size); off += size
    bad = bytearray(good)
    bad[30 + sum(sizes[:5]) + 40] ^= 0x20                       # block 5: CRC mismatch
    bad[30 + sum(sizes[:12]) + 40] ^= 0x20                      # block 12 too: the earlier one decides
    rc1, out1 = gpu.decode(bytes(bad), return_code=True, fill=-7)
    with _Gpus(4):
        rc4, out4 = gpu.decode(bytes(bad), return_code=True, fill=-7)
    assert rc1 == rc4 == harness.DATA_CORRUPTION
    n_ok = 5 * 2048
    assert np.array_equal(out4[:, :n_ok], pcm[:, :n_ok]) and np.array_equal(out1[:, :n_ok], pcm[:, :n_ok])


def test_sharded_packed_pcm_entry_points(gpu, oracle):
    """LINNEB200_EncodeWholePacked / DecodeWholePacked (what the command-line tool calls) take the same block ranges"""
    pcm = harness.synth_pcm(n=4096 * 13 + 900, channels=2, bits=16, seed=41)
    packed = harness.pack_pcm(pcm, 16)
    single = gpu.encode_packed(packed, 2, bits=16, block=4096, preset=2)
    with _Gpus(3):
        sharded = gpu.encode_packed(packed, 2, bits=16, block=4096, preset=2)
        back = gpu.decode_packed(single)
    assert sharded == single
    assert bytes(back) == bytes(packed)
    assert np.array_equal(oracle.decode(sharded), pcm)
