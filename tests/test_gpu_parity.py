"""GPU parity tests: the CUDA path, called through the C-ABI (liblinne_b200.so), against the oracle
(and the unmodified reference where oracle/_ref travelled with the snapshot).

Bar (BASELINE.json north_star): decode bit-exact; encoder output decodes losslessly with the
reference decoder; compressed size within 0.1 % of the reference's at every preset.
"""
import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu

SIZE_TOLERANCE = 0.001      # 0.1 % of the reference's compressed size (north_star)


@pytest.fixture(scope="module")
def clip():
    return harness.synth_pcm(seconds=3.0, channels=2, bits=16, seed=1)


@pytest.mark.parametrize("preset", range(8))
def test_decode_bit_exact(gpu, oracle, clip, preset):
    stream = oracle.encode(clip, preset=preset)
    assert np.array_equal(gpu.decode(stream), clip)


@pytest.mark.parametrize("preset", range(8))
def test_encode_reference_decodable_and_size(gpu, oracle, clip, preset):
    want = oracle.encode(clip, preset=preset)
    got = gpu.encode(clip, preset=preset)
    assert np.array_equal(oracle.decode(got), clip)
    assert abs(len(got) - len(want)) <= SIZE_TOLERANCE * len(want)


def test_encode_byte_identical_on_full_blocks(gpu, oracle):
    # exact-FP build: every full block equals the oracle's bytes (tail blocks with odd unit lengths
    # may differ through the reference's stale-window quirk)
    pcm = harness.synth_pcm(n=10240 * 3, channels=2, bits=16, seed=5)
    for preset in (0, 4, 7):
        assert gpu.encode(pcm, preset=preset) == oracle.encode(pcm, preset=preset)


def test_reference_decodes_gpu_stream(gpu, ref, clip):
    for preset in (0, 7):
        assert np.array_equal(ref.decode(gpu.encode(clip, preset=preset)), clip)


def test_mixed_block_types(gpu, oracle):
    pcm = harness.mixed_types_pcm()
    want = oracle.encode(pcm, preset=7)
    types = []
    off = 30
    while off < len(want):
        types.append(want[off + 8]); off += int.from_bytes(want[off + 2:off + 6], "big") + 6
    assert {0, 1, 2} <= set(types)
    assert np.array_equal(gpu.decode(want), pcm)
    got = gpu.encode(pcm, preset=7)
    assert np.array_equal(oracle.decode(got), pcm)


@pytest.mark.parametrize("channels,bits", [(1, 8), (1, 16), (2, 24), (8, 24), (8, 16)])
def test_formats_roundtrip(gpu, oracle, channels, bits):
    pcm = harness.synth_pcm(n=4096 * 2 + 1500, channels=channels, bits=bits, seed=11)
    for preset in (0, 7):
        got = gpu.encode(pcm, bits=bits, preset=preset, block=4096)
        assert np.array_equal(oracle.decode(got), pcm)
        assert np.array_equal(gpu.decode(oracle.encode(pcm, bits=bits, preset=preset, block=4096)), pcm)


def test_crc_corruption_detected(gpu, oracle, clip):
    stream = bytearray(oracle.encode(clip[:, :30000], preset=0))
    stream[-1] ^= 0x01
    rc, _ = gpu.decode(bytes(stream), return_code=True)
    assert rc == harness.DATA_CORRUPTION
