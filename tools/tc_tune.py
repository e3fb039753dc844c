"""Times k_sweep_tc (forced full-length tensor-core sweep) on the C2 shape; TRACS_TC_DEBUG selects probes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tracs_b200
n, L = 10000, 5_000_000
pitch = (L + 127) // 128 * 128
seqs = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
tracs_b200.synth_device(seqs.data_ptr(), n, L, pitch, seed=2, p_var=0.01, n_clusters=100, mu=5.0, p_N=1e-3, gc=0.508)
for mode in ("tc", True):
    for _ in range(3):
        r = tracs_b200.pairsnp_device(seqs.data_ptr(), n, L, pitch, dist=20, full_sweep=mode, copy=False)
        st = tracs_b200.last_stats()
    print("variant", os.environ.get("TRACS_TC_DEBUG", "0"), "mode", mode, "ms_sweep %.3f" % st["ms_sweep"], "edges", len(r["rows"]))
