#!/usr/bin/env python
"""tools/prof_tput.py -- decode of a long tiled -m PRESET stream (throughput decoder), for ncu captures and stage timings."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import harness  # noqa: E402
from linne_b200 import Product, DecoderSession  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", type=int, default=7)
    ap.add_argument("--blocks", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    codec = Product()
    pcm = harness.synth_pcm(n=10240 * 32, channels=2, bits=16, seed=1)
    base = codec.encode(pcm, preset=args.preset)
    times = max(1, args.blocks // 32)
    stream = harness.tile_stream(base, times, 10240)
    n = pcm.shape[1] * times
    dec = DecoderSession(channels=2)
    out = np.zeros((2, n), np.int32)
    buf = np.frombuffer(stream, np.uint8)
    dec.decode_whole(buf.ctypes.data, len(stream), harness._chan_ptrs(out), 2, n)
    dec.set_profiling(True); dec.reset_stage_stats()
    for _ in range(args.reps):
        dec.decode_whole(buf.ctypes.data, len(stream), harness._chan_ptrs(out), 2, n)
    st = dec.stage_stats()
    assert np.array_equal(out, np.tile(pcm, (1, times))), "decode differs"
    print(f"ok preset={args.preset} blocks={32 * times} " + " ".join(f"{k}={v[1] / v[0]:.3f}ms" for k, v in st.items()))


if __name__ == "__main__":
    main()
