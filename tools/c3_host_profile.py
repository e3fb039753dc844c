"""Where the host time of one C3 call goes: raw C-ABI call vs wrapping vs freeing (diagnosis, not a bench)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import bench as B
import tracs_b200
from tracs_b200 import _lib, api
w = dict(B.CONFIGS["C3"])
if len(sys.argv) > 1:
    w["n"] = int(sys.argv[1]); w["n_clusters"] = max(2, w["n_clusters"] * w["n"] // 100000)
inp = B.Input(torch, tracs_b200, torch.device("cuda"), w)
kw = dict(dist=w["dist"], days=inp.days, lamb=w["lamb"], beta=w["beta"], threshold_Ek=w["threshold_Ek"])
prev = None
for i in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o, keep = api.make_opts(**kw)
    e = _lib.Edges()
    t1 = time.perf_counter()
    _lib.check(_lib.lib().tracs_pairsnp_packed(C.c_void_p(inp.buf.data_ptr()), w["n"], w["L"], inp.pitch, C.byref(o), C.byref(e)))
    t2 = time.perf_counter()
    res = _lib.take_edges(e, names=False, copy=False)
    t3 = time.perf_counter()
    prev = res          # frees the result before last
    t4 = time.perf_counter()
    st = tracs_b200.last_stats()
    print("opts %.3f  call %.3f (device ms_total %.3f)  wrap %.3f  free-previous %.3f ms; edges %d" %
          (1e3 * (t1 - t0), 1e3 * (t2 - t1), st["ms_total"], 1e3 * (t3 - t2), 1e3 * (t4 - t3), len(res["rows"])))
