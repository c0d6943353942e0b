import sys, time
sys.path.insert(0, 'tests')
import numpy as np
import harness
from linne_b200 import Product
g = Product(); o = harness.Oracle()
pcm = harness.synth_pcm(10.0)
for m in range(8):
    t = time.time(); s = g.encode(pcm, preset=m); te = time.time() - t
    t = time.time(); s2 = g.encode(pcm, preset=m); te2 = time.time() - t
    t = time.time(); d = g.decode(s); td = time.time() - t
    t = time.time(); d = g.decode(s); td2 = time.time() - t
    print(f"preset {m}: {len(s)} bytes enc {te:.3f}/{te2:.3f}s dec {td:.3f}/{td2:.3f}s ok={np.array_equal(d, pcm)}", flush=True)
