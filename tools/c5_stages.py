"""tools/c5_stages.py -- per-kernel device times of one six-minute stereo file (a C5 corpus file): encode + decode at -m 0 and -m 7."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, ctypes as C, torch
import harness
from linne_b200 import EncoderSession, DecoderSession
clip = harness.synth_pcm(seconds=10.0, sr=44100, channels=2, bits=16, seed=1)
n = 360 * 44100
pcm = np.ascontiguousarray(np.tile(clip, (1, 36))[:, :n])
stride = (n + 7) // 4 * 4
dev = torch.device('cuda', 0)
d_pcm = torch.zeros((2, stride), dtype=torch.int32, device=dev); d_pcm[:, :n].copy_(torch.from_numpy(pcm))
cap = 30 + 2 * n * 2 + 11 * (n // 10240 + 2) + 65536
d_s = torch.zeros(cap + 64, dtype=torch.uint8, device=dev)
d_b = torch.zeros((2, stride), dtype=torch.int32, device=dev)
for m in (0, 7):
    enc = EncoderSession(2, preset=m); dec = DecoderSession(channels=2)
    sz = enc.encode_whole_resident(d_pcm.data_ptr(), stride, n, d_s.data_ptr(), cap)
    dec.decode_whole_resident(None, d_s.data_ptr(), sz, d_b.data_ptr(), stride, 2, n)
    enc.set_profiling(True); dec.set_profiling(True); enc.reset_stage_stats(); dec.reset_stage_stats()
    for _ in range(3):
        sz = enc.encode_whole_resident(d_pcm.data_ptr(), stride, n, d_s.data_ptr(), cap)
        dec.decode_whole_resident(None, d_s.data_ptr(), sz, d_b.data_ptr(), stride, 2, n)
    print(m, 'enc', {k: round(v[1] / 3, 3) for k, v in enc.stage_stats().items()})
    print(m, 'dec', {k: round(v[1] / 3, 3) for k, v in dec.stage_stats().items()})
    enc.close(); dec.close()
