#!/usr/bin/env python
"""tools/timeline.py -- merged kernel timeline of ONE concurrent -m 0..7 sweep through the host API
(eight presets on eight host threads / streams), from the per-launch CUDA events of every handle."""
import ctypes as C
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch  # noqa: E402
import harness  # noqa: E402
from linne_b200 import EncoderSession, DecoderSession  # noqa: E402

pcm = harness.synth_pcm(seconds=10.0, channels=2, bits=16, seed=1)
nch, n = pcm.shape
cap = 30 + 2 * nch * n * 4 + 65536
h_pcm = torch.from_numpy(pcm.copy()).pin_memory()
encs = {m: EncoderSession(nch, preset=m) for m in range(8)}
decs = {m: DecoderSession(channels=nch) for m in range(8)}
h_outs = {m: torch.zeros(cap, dtype=torch.uint8).pin_memory() for m in range(8)}
h_backs = {m: torch.zeros((nch, n), dtype=torch.int32).pin_memory() for m in range(8)}
chan_in = (C.POINTER(C.c_int32) * nch)(*[C.cast(h_pcm[c].data_ptr(), C.POINTER(C.c_int32)) for c in range(nch)])
chan_outs = {m: (C.POINTER(C.c_int32) * nch)(*[C.cast(h_backs[m][c].data_ptr(), C.POINTER(C.c_int32)) for c in range(nch)]) for m in range(8)}
host = {}


def one(m):
    t0 = time.perf_counter()
    sz = encs[m].encode_whole(chan_in, n, h_outs[m].data_ptr(), cap)
    t1 = time.perf_counter()
    decs[m].decode_whole(h_outs[m].data_ptr(), sz, chan_outs[m], nch, n)
    host[m] = (t0, t1, time.perf_counter())


def sweep():
    th = [threading.Thread(target=one, args=(m,)) for m in range(7, -1, -1)]
    for t in th: t.start()
    for t in th: t.join()


for _ in range(3):
    sweep()
for s in list(encs.values()) + list(decs.values()):
    s.set_profiling(True); s.reset_stage_stats()
torch.cuda.synchronize()
w0 = time.perf_counter()
sweep()
w1 = time.perf_counter()
rows = []
for m in range(8):
    for name, b, e in encs[m].timeline(): rows.append((b, e, m, "enc", name))
    for name, b, e in decs[m].timeline(): rows.append((b, e, m, "dec", name))
t0 = min(r[0] for r in rows)
print(f"sweep wall {1e3 * (w1 - w0):.2f} ms; kernels span {max(r[1] for r in rows) - t0:.2f} ms")
for m in range(8):
    a, b, c = host[m]
    print(f"  m{m}: host encode {1e3 * (a - w0):.2f}..{1e3 * (b - w0):.2f}  decode ..{1e3 * (c - w0):.2f} ms")
for b, e, m, side, name in sorted(rows):
    print(f"{b - t0:8.3f} {e - t0:8.3f}  {e - b:7.3f}  m{m} {side} {name}")
