import os, sys, time, cProfile, pstats
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import bench as B
import tracs_b200
w = B.CONFIGS["C4"]
rng = np.random.default_rng(w["seed"])
msas = []
for r in range(6):
    subset = np.sort(rng.choice(w["n"], size=int(round(rng.uniform(0.6, 1.0) * w["n"])), replace=False))
    L_r = int(rng.integers(2_000_000, 3_000_001))
    wr = dict(w, n=len(subset), L=L_r, seed=w["seed"] + r, n_clusters=max(2, w["n_clusters"] * len(subset) // w["n"]))
    msas.append((subset.astype(np.uint64), B.Input(torch, tracs_b200, torch.device("cuda"), wr), wr))
def step():
    a, b, v = [], [], []
    for subset, inp, wr in msas:
        t0 = time.perf_counter()
        res = tracs_b200.pairsnp_device(inp.buf.data_ptr(), wr["n"], wr["L"], inp.pitch, copy=False, dist=w["dist"])
        t1 = time.perf_counter()
        st = tracs_b200.last_stats()
        a.append(subset[res["rows"].astype(np.int64)]); b.append(subset[res["cols"].astype(np.int64)]); v.append(res["dist"].astype(np.float64))
        print("  msa n=%d L=%d call %.1f ms device %.1f ms edges %d refine %.1f pack %.1f compact %.1f ncomp %.1f sweep %.1f" % (wr["n"], wr["L"], 1e3*(t1-t0), st["ms_total"], len(res["rows"]), st["ms_refine"], st["ms_pack"], st["ms_compact"], st["ms_ncomp"], st["ms_sweep"]))
    t0 = time.perf_counter()
    out = tracs_b200.min_over_refs(np.concatenate(a), np.concatenate(b), np.concatenate(v))
    print("  min_over_refs %.1f ms" % (1e3 * (time.perf_counter() - t0)))
for i in range(3):
    t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print("step %.1f ms" % (1e3 * (time.perf_counter() - t0)))
