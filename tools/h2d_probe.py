import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import tracs_b200
n, L = 2000, 5_000_000
rng = np.random.default_rng(1)
row = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)]
s = np.empty((n, L), np.uint8)
for i in range(n):
    s[i] = row
    s[i, rng.integers(0, L, 50)] = ord("A")
for mode, v in (("staged", "0"), ("driver", str(1 << 40)), ("staged", "0"), ("driver", str(1 << 40))):
    os.environ["TRACS_H2D_STAGE_MIN"] = v
    t = time.time(); r = tracs_b200.pairsnp_matrix(s, dist=100); dt = time.time() - t
    print(mode, "10 GB pageable:", round(dt, 3), "s", len(r["rows"]), "edges", tracs_b200.last_stats()["ms_total"])
