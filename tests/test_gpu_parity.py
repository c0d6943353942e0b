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


# ---- the reference's own round-trip matrix: 9 generators x {1,2,8} ch x {8,16,24} bit x preset {0,7} ----
# test/linne_encode_decode/main.cpp:341-521 (8192 samples, block 1024, M/S for >= 2 channels)
@pytest.mark.parametrize("gen_name", sorted(harness.reference_test_generators()))
def test_reference_roundtrip_matrix(gpu, oracle, gen_name):
    gen = harness.reference_test_generators()[gen_name]
    for channels in (1, 2, 8):
        for bits in (8, 16, 24):
            pcm = harness.to_fixed(gen(channels, 8192), bits)
            for preset in (0, 7):
                stream = gpu.encode(pcm, bits=bits, rate=8000, block=1024, preset=preset)
                assert np.array_equal(gpu.decode(stream), pcm), (gen_name, channels, bits, preset)
                assert np.array_equal(oracle.decode(stream), pcm), (gen_name, channels, bits, preset)


@pytest.mark.parametrize("name", harness.GOLDEN_CASES + ["mixed_types_m5"])
def test_golden_decode_bit_exact(gpu, name):
    g = harness.load_golden(name)
    assert np.array_equal(gpu.decode(g["stream"].tobytes()), g["pcm"])


@pytest.mark.parametrize("name", harness.GOLDEN_CASES)
def test_golden_coefficients_give_identical_bytes(gpu, name):
    # north star leg 3: identical quantised coefficients -> identical residuals and coded bits
    g = harness.load_golden(name)
    got = gpu.encode_with_params(g["pcm"], harness.params_from_golden(g), bits=int(g["bits"]),
                                 block=int(g["block"]), preset=int(g["preset"]))
    assert got == g["stream"].tobytes()


def test_block_sizes_on_both_kernel_paths(gpu, oracle):
    # 2048/4096/8192/10240 take the cooperative kernels, 1024/3000/12000 the flat ones; all must agree with the oracle
    for block in (1024, 2048, 3000, 4096, 8192, 10240, 12000):
        pcm = harness.synth_pcm(n=block * 2 + block // 3, channels=2, bits=16, seed=block)
        for preset in (1, 6):
            got = gpu.encode(pcm, preset=preset, block=block)
            want = oracle.encode(pcm, preset=preset, block=block)
            assert np.array_equal(oracle.decode(got), pcm), (block, preset)
            assert np.array_equal(gpu.decode(want), pcm), (block, preset)
            assert abs(len(got) - len(want)) <= max(2, SIZE_TOLERANCE * len(want)), (block, preset, len(got), len(want))


def test_config_c4_24bit_96k_8ch(gpu, oracle):
    # BASELINE.json configs[3]: 24-bit / 96 kHz / 8 channels at -m 7 (long-order predictor, multichannel path)
    pcm = harness.synth_pcm(seconds=1.0, sr=96000, channels=8, bits=24, seed=44)
    got = gpu.encode(pcm, bits=24, rate=96000, preset=7)
    want = oracle.encode(pcm, bits=24, rate=96000, preset=7)
    assert np.array_equal(oracle.decode(got), pcm)
    assert np.array_equal(gpu.decode(want), pcm)
    assert abs(len(got) - len(want)) <= SIZE_TOLERANCE * len(want)


def test_config_c3_long_stream_decode(gpu, oracle):
    # BASELINE.json configs[2] shape: a long -m 7 stream built by tiling the blocks of a verified short one
    # (legal because blocks are self-contained); property: decode == tiled PCM
    pcm = harness.synth_pcm(n=10240 * 12, channels=2, bits=16, seed=55)
    base = gpu.encode(pcm, preset=7)
    assert np.array_equal(oracle.decode(base), pcm)
    times = 40                                     # 480 blocks, ~111 s of audio
    long_stream = harness.tile_stream(base, times, 10240)
    out = gpu.decode(long_stream)
    assert out.shape[1] == pcm.shape[1] * times
    assert np.array_equal(out, np.tile(pcm, (1, times)))


def test_config_c5_block_range_shards(gpu):
    # BASELINE.json configs[4] property: encoding contiguous block ranges separately and concatenating them
    # equals the single-call stream byte for byte, so ranks can take block ranges with no exchange but the gather
    from linne_b200 import shard
    pcm = harness.synth_pcm(n=10240 * 9 + 4000, channels=2, bits=16, seed=66)
    whole = gpu.encode(pcm, preset=4)
    for world in (2, 4, 8):
        parts = [shard.encode_shard(gpu, pcm, r, world, 10240, preset=4) for r in range(world)]
        assert shard.assemble(parts[0][0], [p[1] for p in parts]) == whole
        out = np.zeros_like(pcm)
        for r in range(world):
            first, got = shard.decode_shard(gpu, whole, r, world)
            if got is not None:
                out[:, first:first + got.shape[1]] = got
        assert np.array_equal(out, pcm)


# ---- non-default analysis paths: CLI -a N and -l ----
@pytest.mark.parametrize("preset,block", [(0, 4096), (2, 4096), (5, 1024)])
def test_irls_refinement_path(gpu, oracle, preset, block):
    pcm = harness.synth_pcm(n=block * 2 + block // 2, channels=2, bits=16, seed=91 + preset)
    got = gpu.encode(pcm, preset=preset, block=block, af=2)
    want = oracle.encode(pcm, preset=preset, block=block, af=2)
    assert np.array_equal(oracle.decode(got), pcm)
    assert abs(len(got) - len(want)) <= max(2, SIZE_TOLERANCE * len(want)), (len(got), len(want))
    plain = gpu.encode(pcm, preset=preset, block=block)
    assert got != plain                      # the refinement really ran


def _blocks(stream):
    """[(offset, size)] of the blocks of a stream"""
    out, off = [], 30
    while off < len(stream):
        size = int.from_bytes(stream[off + 2:off + 6], "big") + 6
        out.append((off, size)); off += size
    return out


@pytest.mark.parametrize("preset", range(8))
def test_c2_clip_exact_shape_per_mode(gpu, oracle, preset):
    """BASELINE.json configs[1] at its real size: the 10 s clip (43 full blocks + a 680-sample tail), every mode.
    Size within 0.1 % of the reference's, lossless through the reference decoder's restatement, and -- so that a
    regression in full blocks cannot hide inside the size bound -- every FULL block byte-identical; only the tail
    block (odd unit lengths: the reference's stale window sample) may differ."""
    pcm = harness.synth_pcm(seconds=10.0, channels=2, bits=16, seed=1)
    want = oracle.encode(pcm, preset=preset)
    got = gpu.encode(pcm, preset=preset)
    assert np.array_equal(oracle.decode(got), pcm)
    assert np.array_equal(gpu.decode(want), pcm)
    assert abs(len(got) - len(want)) <= SIZE_TOLERANCE * len(want), (len(got), len(want))
    bw, bg = _blocks(want), _blocks(got)
    assert len(bw) == len(bg) == 44
    differing = [i for i, ((ow, sw), (og, sg)) in enumerate(zip(bw, bg)) if want[ow:ow + sw] != got[og:og + sg]]
    assert differing in ([], [43]), differing


@pytest.mark.parametrize("preset,block", [(0, 2048), (0, 10240), (7, 10240)])
def test_sgd_training_path(gpu, oracle, preset, block):
    """enable_learning: 2000 momentum-SGD steps on the L1 loss (linne_network.c:805-873).  One block of the CLI's size at
    the cheapest and the most expensive preset: lossless, and the size within the north star's 0.1 % of the reference's
    (the summation order of the gradient differs from the CPU's; the 8-bit quantiser absorbs it)."""
    pcm = harness.synth_pcm(n=block, channels=2, bits=16, seed=97 + preset)
    got = gpu.encode(pcm, preset=preset, block=block, learning=1)
    want = oracle.encode(pcm, preset=preset, block=block, learning=1)
    assert np.array_equal(oracle.decode(got), pcm)
    assert abs(len(got) - len(want)) <= max(4, SIZE_TOLERANCE * len(want)), (len(got), len(want))
    assert got != gpu.encode(pcm, preset=preset, block=block)


@pytest.mark.parametrize("kind", ["af", "learning"])
def test_refinement_on_blocks_longer_than_shared_memory(gpu, oracle, kind):
    """IRLS / SGD on 16384-sample blocks (the reference CLI's capacity, linne_codec.c:54): the refinement kernel keeps its
    signals in a global scratch there; same bounds as for short blocks."""
    block = 16384
    pcm = harness.synth_pcm(n=block + block // 2, channels=2, bits=16, seed=131)
    opts = {"af": 2} if kind == "af" else {"learning": 1}
    got = gpu.encode(pcm, preset=0, block=block, **opts)
    want = oracle.encode(pcm, preset=0, block=block, **opts)
    assert np.array_equal(oracle.decode(got), pcm)
    assert abs(len(got) - len(want)) <= max(4, SIZE_TOLERANCE * len(want)), (len(got), len(want))
    assert got != gpu.encode(pcm, preset=0, block=block)


# ---- robustness: corrupted payloads with the CRC check off must come back (any result code), never hang ----
@pytest.mark.timeout(120)
def test_corrupted_payloads_do_not_hang(gpu, oracle):
    rng = np.random.default_rng(1234)
    pcm = harness.synth_pcm(n=4096 * 3 + 700, channels=2, bits=16, seed=55)
    for preset in (0, 5):
        good = oracle.encode(pcm, preset=preset, block=4096)
        for trial in range(24):
            bad = bytearray(good)
            # leave the 30-byte stream header and the block framing (sync, size) alone: hit type / payload bytes
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(41, len(bad)))
                bad[pos] = int(rng.integers(0, 256))
            rc, out = gpu.decode(bytes(bad), check_crc=0, return_code=True)
            assert rc in range(8)
    # and the undamaged stream still decodes afterwards (no state left behind)
    assert np.array_equal(gpu.decode(good), pcm)
