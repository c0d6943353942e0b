"""GPU parity tests of the throughput decoder (linne_b200/csrc/lnb_tput_v2.cuh: eight lanes per block / one lane per
(block, channel), taken by large batches).  LINNE_B200_TPUT_MIN_BLOCKS moves the switch-over point, so the same
small streams run through it (=1) and through the per-block pipeline (=0); both must equal the oracle's PCM bit for
bit, and result codes and partial outputs of damaged streams must be the same on both paths."""
import os

import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu


@pytest.fixture
def tput():
    """switch the throughput path on for every batch"""
    old = os.environ.get("LINNE_B200_TPUT_MIN_BLOCKS")
    os.environ["LINNE_B200_TPUT_MIN_BLOCKS"] = "1"
    yield
    if old is None:
        del os.environ["LINNE_B200_TPUT_MIN_BLOCKS"]
    else:
        os.environ["LINNE_B200_TPUT_MIN_BLOCKS"] = old


def _without(fn):
    old = os.environ.get("LINNE_B200_TPUT_MIN_BLOCKS")
    os.environ["LINNE_B200_TPUT_MIN_BLOCKS"] = "0"
    try:
        return fn()
    finally:
        if old is None:
            del os.environ["LINNE_B200_TPUT_MIN_BLOCKS"]
        else:
            os.environ["LINNE_B200_TPUT_MIN_BLOCKS"] = old


def _stages(gpu, stream, channels):
    """kernel names the decoder launched for this stream"""
    import ctypes as C
    from linne_b200 import DecoderSession
    dec = DecoderSession(channels=channels)
    dec.set_profiling(True)
    rc, hdr = gpu.decode_header(stream)
    out = np.zeros((hdr.num_channels, hdr.num_samples), np.int32)
    buf = np.frombuffer(stream, np.uint8)
    dec.decode_whole(buf.ctypes.data, len(stream), harness._chan_ptrs(out), hdr.num_channels, hdr.num_samples)
    names = set(dec.stage_stats())
    dec.close()
    return names, out


@pytest.mark.parametrize("preset", range(8))
def test_tput_decode_bit_exact_every_preset(gpu, oracle, tput, preset):
    pcm = harness.synth_pcm(n=10240 * 5 + 3000, channels=2, bits=16, seed=70 + preset)      # 5 full blocks + a tail block
    stream = oracle.encode(pcm, preset=preset)
    names, out = _stages(gpu, stream, 2)
    assert {"tp_entropy", "tp_synth"} <= names, names          # the path under test really ran
    assert np.array_equal(out, pcm)


@pytest.mark.parametrize("channels,bits,block", [(1, 8, 1024), (1, 16, 4096), (2, 24, 2048), (8, 24, 10240), (8, 16, 3072), (4, 16, 1024)])
def test_tput_formats(gpu, oracle, tput, channels, bits, block):
    pcm = harness.synth_pcm(n=block * 4 + block // 3, channels=channels, bits=bits, seed=17 + channels + bits)
    for preset in (1, 3, 7):
        stream = oracle.encode(pcm, bits=bits, preset=preset, block=block)
        assert np.array_equal(gpu.decode(stream), pcm), (preset,)


def test_tput_many_blocks_unit_counts_and_types(gpu, oracle, tput):
    """many blocks (several warps of lanes), signals that make the encoder choose different unit counts per block
    and layer, raw and silent blocks in between"""
    rng = np.random.default_rng(5)
    n = 2048
    parts = []
    for i in range(70):
        kind = i % 7
        t = np.arange(n)
        if kind == 0:
            x = 8000 * np.sin(2 * np.pi * (200 + 37 * i) * t / 44100.0)
        elif kind == 1:
            x = rng.normal(0, 3000, n)                                          # noise: raw or short predictors
        elif kind == 2:
            x = np.zeros(n)                                                     # silent block
        elif kind == 3:
            x = np.concatenate([4000 * np.sin(2 * np.pi * f * t[:n // 8] / 44100.0) for f in rng.integers(100, 9000, 8)])  # changes every unit
        elif kind == 4:
            x = np.cumsum(rng.normal(0, 40, n))
        elif kind == 5:
            x = rng.integers(-32768, 32767, n).astype(float)                    # full-scale noise: raw block
        else:
            x = 12000 * np.sign(np.sin(2 * np.pi * 440 * t / 44100.0)) + rng.normal(0, 20, n)
        parts.append(np.stack([x, np.roll(x, 5) * 0.7 + rng.normal(0, 10, n)]))
    pcm = np.clip(np.concatenate(parts, axis=1), -32768, 32767).astype(np.int32)
    for preset in (0, 4, 7):
        stream = oracle.encode(pcm, preset=preset, block=n)
        assert np.array_equal(gpu.decode(stream), pcm), preset
        assert np.array_equal(_without(lambda: gpu.decode(stream)), pcm), preset


def test_tput_long_stream_default_threshold(gpu, oracle):
    """above the default switch-over point (2560 blocks) without any override: 4200 blocks of 1024 samples"""
    pcm = harness.synth_pcm(n=1024 * 20, channels=2, bits=16, seed=31)
    base = oracle.encode(pcm, preset=6, block=1024)
    times = 210
    stream = harness.tile_stream(base, times, 1024)
    names, out = _stages(gpu, stream, 2)
    assert {"tp_entropy", "tp_synth"} <= names, names
    assert np.array_equal(out, np.tile(pcm, (1, times)))


@pytest.mark.timeout(180)
def test_tput_damaged_streams_same_as_per_block_path(gpu, oracle, tput):
    rng = np.random.default_rng(99)
    pcm = harness.synth_pcm(n=2048 * 6 + 500, channels=2, bits=16, seed=56)
    for preset in (0, 5):
        good = oracle.encode(pcm, preset=preset, block=2048)
        for trial in range(20):
            bad = bytearray(good)
            for _ in range(int(rng.integers(1, 5))):
                pos = int(rng.integers(41, len(bad)))
                bad[pos] = int(rng.integers(0, 256))
            for crc in (1, 0):
                rc, out = gpu.decode(bytes(bad), check_crc=crc, return_code=True, fill=-7)
                rc2, out2 = _without(lambda: gpu.decode(bytes(bad), check_crc=crc, return_code=True, fill=-7))
                assert rc == rc2, (preset, trial, crc)
                if crc:                     # with the CRC check on, the delivered samples are defined by the reference
                    assert np.array_equal(out, out2), (preset, trial)
        assert np.array_equal(gpu.decode(good), pcm)


@pytest.fixture
def env_switch():
    """set environment switches the library reads per call, restored afterwards"""
    saved = {}

    def setenv(name, value):
        saved.setdefault(name, os.environ.get(name))
        os.environ[name] = value
    yield setenv
    for name, old in saved.items():
        if old is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = old


@pytest.mark.parametrize("form", ["i", "8", "w", "l", "s"])
def test_tput_entropy_forms_bit_exact(gpu, oracle, tput, env_switch, form):
    """every form of the throughput entropy stage (default: fix-up rounds at eight lanes per block; the measured
    alternatives stay selectable) decodes the same streams to the oracle's PCM, including damaged ones"""
    env_switch("LINNE_B200_TP_ENTROPY", form)
    rng = np.random.default_rng(7)
    for channels, bits, block, preset in [(2, 16, 10240, 7), (2, 16, 1024, 0), (8, 24, 2048, 5), (1, 8, 1024, 2)]:
        pcm = harness.synth_pcm(n=block * 6 + block // 2, channels=channels, bits=bits, seed=3 + channels + preset)
        if channels == 2 and block == 1024:
            pcm[:, 700:712] = np.array([30000, -30000] * 6, np.int32)        # a click in a quiet partition: code words longer than 32 bits
        stream = oracle.encode(pcm, bits=bits, preset=preset, block=block)
        names, out = _stages(gpu, stream, channels)
        assert {"tp_entropy", "tp_synth"} <= names, names
        assert np.array_equal(out, pcm), (form, channels, bits, block, preset)
    pcm = harness.synth_pcm(n=2048 * 6 + 500, channels=2, bits=16, seed=56)
    good = oracle.encode(pcm, preset=3, block=2048)
    for trial in range(8):
        bad = bytearray(good)
        for _ in range(int(rng.integers(1, 5))):
            bad[int(rng.integers(41, len(bad)))] = int(rng.integers(0, 256))
        rc, out = gpu.decode(bytes(bad), check_crc=1, return_code=True, fill=-7)
        rc2, out2 = _without(lambda: gpu.decode(bytes(bad), check_crc=1, return_code=True, fill=-7))
        assert rc == rc2 and np.array_equal(out, out2), (form, trial)


@pytest.mark.parametrize("walk", ["lane", "warp"])
def test_fused_decoder_walk_forms_bit_exact(gpu, oracle, env_switch, walk):
    """the fused per-block decoder with its walk on one lane (default) and as a warp (rounds of 32 code words)"""
    env_switch("LINNE_B200_WALK", walk)
    env_switch("LINNE_B200_TPUT_MIN_BLOCKS", "0")
    for channels, bits, block, preset in [(2, 16, 10240, 7), (2, 16, 4096, 0), (8, 24, 2048, 4), (1, 8, 1024, 2)]:
        pcm = harness.synth_pcm(n=block * 3 + block // 3, channels=channels, bits=bits, seed=11 + channels + preset)
        if channels == 2 and block == 4096:
            pcm[:, 900:906] = np.array([32000, -32000] * 3, np.int32)
        stream = oracle.encode(pcm, bits=bits, preset=preset, block=block)
        assert np.array_equal(gpu.decode(stream), pcm), (walk, channels, bits, block, preset)
