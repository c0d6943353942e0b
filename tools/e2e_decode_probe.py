#!/usr/bin/env python
"""tools/e2e_decode_probe.py -- wall time of LINNEDecoder_DecodeWhole / LINNEB200_DecodeWholePacked with page-locked host
buffers on a long tiled -m PRESET stream, for several settings of LINNE_B200_PIPELINE (block ranges per device)."""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import harness  # noqa: E402
from linne_b200 import Product, DecoderSession  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", type=int, default=7)
    ap.add_argument("--blocks", type=int, default=15504)
    ap.add_argument("--depths", default="1,2,4,8")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    codec = Product()
    pcm = harness.synth_pcm(n=10240 * 32, channels=2, bits=16, seed=1)
    base = codec.encode(pcm, preset=args.preset)
    times = max(1, args.blocks // 32)
    stream = harness.tile_stream(base, times, 10240)
    n = pcm.shape[1] * times
    want = np.tile(pcm, (1, times))
    h_stream = torch.zeros(len(stream) + 16, dtype=torch.uint8).pin_memory()
    h_stream[:len(stream)] = torch.from_numpy(np.frombuffer(stream, np.uint8).copy())
    h_back = torch.zeros((2, n), dtype=torch.int32).pin_memory()
    h_packed = torch.zeros(n * 2 * 2, dtype=torch.uint8).pin_memory()
    chan_out = (C.POINTER(C.c_int32) * 2)(*[C.cast(h_back[c].data_ptr(), C.POINTER(C.c_int32)) for c in range(2)])
    u8p = C.POINTER(C.c_uint8)
    dec = DecoderSession(channels=2)
    frames = C.c_uint32(0)
    for depth in [d for d in args.depths.split(",")]:
        if depth == "default":
            os.environ.pop("LINNE_B200_PIPELINE", None)
        else:
            os.environ["LINNE_B200_PIPELINE"] = depth
        res = {}
        for name in ("planes", "packed"):
            best = 1e9
            for _ in range(args.reps + 1):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if name == "planes":
                    dec.decode_whole(h_stream.data_ptr(), len(stream), chan_out, 2, n)
                else:
                    rc = dec.lib.LINNEB200_DecodeWholePacked(dec.h, C.cast(h_stream.data_ptr(), u8p), len(stream),
                                                             C.cast(h_packed.data_ptr(), u8p), n, C.byref(frames))
                    assert rc == 0
                best = min(best, (time.perf_counter() - t0) * 1e3)
            res[name] = best
        ok = bool(np.array_equal(h_back.numpy(), want)) and bool(np.array_equal(h_packed.numpy().view("<i2").reshape(n, 2).T, want))
        print(f"pipeline={depth} blocks={32 * times} planes={res['planes']:.2f}ms packed={res['packed']:.2f}ms ok={ok} "
              f"(up {len(stream) / 1e6:.0f} MB, down {8 * n / 1e6:.0f} / {4 * n / 1e6:.0f} MB)", flush=True)
    dec.close()


if __name__ == "__main__":
    main()
