#!/usr/bin/env python
"""tools/prof_decode.py -- per-kernel device times of DecodeWhole on the synthetic clip (per-block pipeline kernel),
one line per preset: python tools/prof_decode.py --presets 0,7 --seconds 10 --reps 5"""
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
    ap.add_argument("--presets", default="0,7")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--bits", type=int, default=16)
    ap.add_argument("--rate", type=int, default=44100)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    codec = Product()
    pcm = harness.synth_pcm(seconds=args.seconds, sr=args.rate, channels=args.channels, bits=args.bits, seed=1)
    C, n = pcm.shape
    for preset in [int(p) for p in args.presets.split(",")]:
        stream = codec.encode(pcm, bits=args.bits, rate=args.rate, preset=preset)
        dec = DecoderSession(channels=C)
        out = np.zeros((C, n), np.int32)
        buf = np.frombuffer(stream, np.uint8)
        dec.decode_whole(buf.ctypes.data, len(stream), harness._chan_ptrs(out), C, n)
        dec.set_profiling(True); dec.reset_stage_stats()
        for _ in range(args.reps):
            dec.decode_whole(buf.ctypes.data, len(stream), harness._chan_ptrs(out), C, n)
        st = dec.stage_stats()
        assert os.environ.get("LINNE_B200_DBG_WALKONLY") or np.array_equal(out, pcm), "decode differs"
        print(f"ok preset={preset} samples={pcm.size} bytes={len(stream)} " +
              " ".join(f"{k}={v[1] / v[0]:.3f}ms" for k, v in st.items()), flush=True)
        dec.close()


if __name__ == "__main__":
    main()
