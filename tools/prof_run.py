#!/usr/bin/env python
"""tools/prof_run.py -- a short encode + decode of the synthetic clip at one preset, for ncu captures
and quick stage timings (python tools/prof_run.py --preset 7 --seconds 10 --reps 2 [--stages])."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import harness  # noqa: E402
from linne_b200 import Product  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", type=int, default=7)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--bits", type=int, default=16)
    ap.add_argument("--rate", type=int, default=44100)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--af", type=int, default=0, help="IRLS iterations (-a)")
    ap.add_argument("--learning", type=int, default=0, help="momentum SGD (-l)")
    args = ap.parse_args()
    codec = Product()
    pcm = harness.synth_pcm(seconds=args.seconds, sr=args.rate, channels=args.channels, bits=args.bits, seed=1)
    for _ in range(args.reps):
        stream = codec.encode(pcm, bits=args.bits, rate=args.rate, preset=args.preset, af=args.af, learning=args.learning)
        back = codec.decode(stream)
    assert np.array_equal(back, pcm), "round trip differs"
    print(f"ok preset={args.preset} samples={pcm.size} bytes={len(stream)}")


if __name__ == "__main__":
    main()
