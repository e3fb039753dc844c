import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import tracs_b200
n, L = 23680, 1_000_000
pitch = L // 2
buf = torch.empty(n * pitch, dtype=torch.uint8, device="cuda")
tracs_b200.synth_device(buf.data_ptr(), n, L, pitch, seed=5, p_var=1.0, n_clusters=236, mu=5.0, p_N=0.0, gc=0.5, gaps=0, packed=True)
for dbg in ("0", "1", "2", "4", "8", "15"):
    os.environ["TRACS_TC3_DBG"] = dbg
    ts = []
    for _ in range(2):
        tracs_b200.pairsnp_packed(buf.data_ptr(), n, L, pitch, dist=20, full_sweep="tc")
        ts.append(tracs_b200.last_stats()["ms_sweep"])
    print("dbg", dbg, "ms_sweep", [round(t, 1) for t in ts], flush=True)
