"""One C4-shaped MSA (1600 samples x 2.5 Mb, 30 % N, 5 % ambiguity codes, dist <= 100) through pairsnp_device: stage
timers, for profiling the N-rich case under ncu."""
import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import tracs_b200
n, L = 1600, 2_500_000
pitch = (L + 127) // 128 * 128
buf = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
tracs_b200.synth_device(buf.data_ptr(), n, L, pitch, seed=4, p_var=0.01, n_clusters=32, mu=5.0, p_N=0.3, p_amb=0.05, gc=0.5, gaps=2)
for rep in range(3):
    r = tracs_b200.pairsnp_device(buf.data_ptr(), n, L, pitch, dist=100)
    st = tracs_b200.last_stats()
    print(rep, len(r["rows"]), {k: round(v, 2) for k, v in st.items() if k.startswith("ms_") and v}, st["n_candidates"], st["swept_wordpairs"] // max(1, st["n_pairs"]), flush=True)
