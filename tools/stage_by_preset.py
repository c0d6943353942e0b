#!/usr/bin/env python
"""tools/stage_by_preset.py -- per-kernel device time (CUDA events) of one encode + decode of the C2 clip at
every preset, run one preset at a time (kernels alone on the GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import harness  # noqa: E402
from linne_b200 import EncoderSession, DecoderSession  # noqa: E402
import ctypes as C  # noqa: E402

pcm = harness.synth_pcm(seconds=10.0, channels=2, bits=16, seed=1)
nch, n = pcm.shape
cap = 30 + 2 * nch * n * 4 + 65536
out = np.zeros(cap, dtype=np.uint8)
back = np.zeros((nch, n), dtype=np.int32)
chan_in = (C.POINTER(C.c_int32) * nch)(*[C.cast(pcm[c].ctypes.data, C.POINTER(C.c_int32)) for c in range(nch)])
chan_out = (C.POINTER(C.c_int32) * nch)(*[C.cast(back[c].ctypes.data, C.POINTER(C.c_int32)) for c in range(nch)])
for m in range(8):
    enc, dec = EncoderSession(nch, preset=m), DecoderSession(channels=nch)
    for rep in range(3):
        if rep == 2:
            enc.set_profiling(True); dec.set_profiling(True); enc.reset_stage_stats(); dec.reset_stage_stats()
        sz = enc.encode_whole(chan_in, n, out.ctypes.data, cap)
        dec.decode_whole(out.ctypes.data, sz, chan_out, nch, n)
    assert np.array_equal(back, pcm)
    st = {**enc.stage_stats(), **dec.stage_stats()}
    print(f"m{m}: " + "  ".join(f"{k}={v[1]:.3f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1][1])[:6]))
    enc.close(); dec.close()
