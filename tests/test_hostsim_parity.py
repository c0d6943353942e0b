"""CPU-only parity of the product's host code + kernel bodies (tests/hostsim: the same C host code
and the same lnb_*_core.cuh work-item functions, compiled for the CPU with a loop executor) against
the oracle and the golden fixtures.  This is what keeps kernel logic honest between GPU runs; the
GPU suite (test_gpu_parity.py) repeats these through liblinne_b200.so on the B200."""
import numpy as np
import pytest

import harness


@pytest.mark.parametrize("name", harness.GOLDEN_CASES + ["mixed_types_m5"])
def test_golden_decode_bit_exact(hostsim, name):
    g = harness.load_golden(name)
    assert np.array_equal(hostsim.decode(g["stream"].tobytes()), g["pcm"])


@pytest.mark.parametrize("name", harness.GOLDEN_CASES)
def test_golden_coefficients_give_identical_bytes(hostsim, name):
    # north star leg 3: identical quantised coefficients -> identical residuals and coded bits
    g = harness.load_golden(name)
    got = hostsim.encode_with_params(g["pcm"], harness.params_from_golden(g), bits=int(g["bits"]),
                                     block=int(g["block"]), preset=int(g["preset"]))
    assert got == g["stream"].tobytes()


@pytest.mark.parametrize("preset", range(8))
def test_encode_matches_oracle_on_full_blocks(hostsim, oracle, preset):
    pcm = harness.synth_pcm(n=4096 * 2, channels=2, bits=16, seed=60 + preset)
    assert hostsim.encode(pcm, preset=preset, block=4096) == oracle.encode(pcm, preset=preset, block=4096)


@pytest.mark.parametrize("preset", (0, 3, 7))
def test_ragged_tail_is_lossless_and_close(hostsim, oracle, preset):
    # tail blocks with odd unit lengths may differ from the reference bytes (stale window sample)
    pcm = harness.synth_pcm(n=4096 + 1737, channels=2, bits=16, seed=70)
    got = hostsim.encode(pcm, preset=preset, block=4096)
    want = oracle.encode(pcm, preset=preset, block=4096)
    assert np.array_equal(oracle.decode(got), pcm)
    assert abs(len(got) - len(want)) <= 0.001 * len(want)


def test_reference_generators_roundtrip(hostsim, oracle):
    # test/linne_encode_decode/main.cpp:341-521 (subset of the 162 cases; the GPU suite runs them all)
    for name, gen in harness.reference_test_generators().items():
        for channels, bits, preset in ((1, 16, 0), (2, 8, 7), (8, 24, 7)):
            pcm = harness.to_fixed(gen(channels, 2048), bits)
            got = hostsim.encode(pcm, bits=bits, rate=8000, block=1024, preset=preset)
            assert np.array_equal(oracle.decode(got), pcm), (name, channels, bits, preset)
            assert np.array_equal(hostsim.decode(got), pcm), (name, channels, bits, preset)


def test_short_tail_blocks_do_not_crash(hostsim, oracle):
    # the reference crashes on tails shorter than ~2*P_max at presets 5-7; we must not
    for tail in (1, 2, 7, 100, 130, 255):
        pcm = harness.synth_pcm(n=1024 + tail, channels=1, bits=16, seed=80 + tail)
        got = hostsim.encode(pcm, preset=7, block=1024)
        assert np.array_equal(hostsim.decode(got), pcm), tail
        assert np.array_equal(oracle.decode(got), pcm), tail
