"""Packed-PCM entry points: LINNEB200_EncodeWholePacked / DecodeWholePacked take and give the
bytes of a WAV data chunk; the conversion runs on the device.  Must equal the int32-planar calls bit for bit."""
import numpy as np
import pytest

import harness
from linne_b200 import OK, DATA_CORRUPTION

CASES = [(2, 16, 3, 10240 + 3000), (1, 8, 0, 5000), (3, 24, 5, 4096 * 2 + 100), (8, 24, 2, 3000)]


def check(codec, oracle):
    for ch, bits, preset, n in CASES:
        pcm = harness.synth_pcm(n=n, channels=ch, bits=bits, seed=70 + ch + bits)
        packed = harness.pack_pcm(pcm, bits)
        planar_stream = codec.encode(pcm, bits=bits, preset=preset, block=4096)
        assert codec.encode_packed(packed, ch, bits=bits, preset=preset, block=4096) == planar_stream
        assert codec.decode_packed(planar_stream) == packed
        assert np.array_equal(oracle.decode(planar_stream), pcm)
    # a corrupted block: everything before it is still handed back, with the reference's result code
    ch, bits, preset, n = CASES[0]
    pcm = harness.synth_pcm(n=n, channels=ch, bits=bits, seed=99)
    stream = bytearray(codec.encode(pcm, bits=bits, preset=preset, block=4096))
    table = []
    off = 30
    while off < len(stream):
        size = int.from_bytes(stream[off + 2:off + 6], "big") + 6
        table.append((off, size)); off += size
    stream[table[2][0] + 40] ^= 0x55
    rc, got = codec.decode_packed(bytes(stream), return_code=True)
    assert rc == DATA_CORRUPTION
    assert got == harness.pack_pcm(pcm[:, :2 * 4096], bits)


def test_packed_pcm_hostsim(hostsim, oracle):
    check(hostsim, oracle)


@pytest.mark.gpu
def test_packed_pcm_gpu(gpu, oracle):
    check(gpu, oracle)
