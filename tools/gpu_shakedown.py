"""First-contact GPU script: INT peak micro-benchmark + a mid-size timing on device-generated data."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tracs_b200
from tracs_b200 import _lib

os.makedirs("gpurun_out", exist_ok=True)
os.system("nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv; nproc; free -g | head -2")
pk = tracs_b200.int_peak()
print("INT peak:", json.dumps(pk))
json.dump(pk, open("gpurun_out/int_peak.json", "w"))

for (n, L, C_, dist) in [(1000, 2_800_000, 20, 20), (4000, 1_000_000, 40, 20), (10000, 5_000_000, 100, 20)]:
    pitch = (L + 127) // 128 * 128
    p = C.c_void_p()
    _lib.check(_lib.lib().tracs_dev_alloc(C.byref(p), n * pitch))
    d = C.c_void_p()
    _lib.check(_lib.lib().tracs_dev_alloc(C.byref(d), n * 4))
    t0 = time.time()
    tracs_b200.synth_device(p.value, n, L, pitch, seed=2, p_var=0.01, n_clusters=C_, mu=5.0, p_N=1e-3, gc=0.5, dev_days=d.value)
    print("synth %dx%d: %.2fs" % (n, L, time.time() - t0))
    days = np.zeros(n, np.int32)
    _lib.check(_lib.lib().tracs_memcpy_d2h(days.ctypes.data, d, n * 4))
    for rep in range(2):
        t0 = time.time()
        res = tracs_b200.pairsnp_device(p.value, n, L, pitch, dist=dist, days=days)
        wall = time.time() - t0
        st = tracs_b200.last_stats()
        P = n * (n - 1) // 2
        print(json.dumps({"n": n, "L": L, "wall_s": round(wall, 3), "site_pairs_per_s_total": P * L / (st["ms_total"] * 1e-3),
                          "site_pairs_per_s_sweep": P * L / (st["ms_sweep"] * 1e-3),
                          "wordpairs_per_s_sweep": P * st["n_words"] / (st["ms_sweep"] * 1e-3), **st}))
    _lib.lib().tracs_dev_free(p)
    _lib.lib().tracs_dev_free(d)
