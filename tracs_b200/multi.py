"""Multi-GPU plumbing for the pair sweep: one process per GPU, each sweeps the row-blocks dealt to its
rank (tracs_opts_t.shard_rank/shard_world; no traffic during the sweep), then ONE variable-length
gather of the edge columns to rank 0, merged back into (row, col) order. torch.distributed is the
transport (NCCL on GPUs; gloo in the CPU tests)."""
import numpy as np

COLUMNS = ("rows", "cols", "dist", "ncomp", "p0_log", "eK")


def merge_sorted(parts):
    """parts: list of float64[6][E_r] blocks, each sorted by (row, col) and with disjoint rows."""
    allp = np.concatenate(parts, axis=1) if parts else np.zeros((len(COLUMNS), 0))
    key = (allp[0].astype(np.uint64) << np.uint64(32)) | allp[1].astype(np.uint64)
    order = np.argsort(key, kind="stable")
    return allp[:, order]


def gather_edges(res, rank, world, dist, torch, device, merge=True):
    """All ranks call this; rank 0 gets the float64[6][E] table (rows, cols, d, ncomp, p0, eK) merged into
    (row, col) order -- or, with merge=False, the list of per-rank tables (one MSA per rank) -- the
    others get None. Integer columns stay exact in float64 (< 2^53)."""
    n_loc = len(res["rows"])
    cnt = torch.tensor([n_loc], dtype=torch.int64, device=device)
    cnts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    mx = max(cnts + [1])
    pack = np.zeros((len(COLUMNS), mx), dtype=np.float64)
    for k, c in enumerate(COLUMNS):
        v = res.get(c)
        if v is not None and n_loc:
            pack[k, :n_loc] = v
    t = torch.from_numpy(pack).to(device)
    if rank == 0:
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.gather(t, bufs, dst=0)
        parts = [b[:, :cnts[r]].cpu().numpy() for r, b in enumerate(bufs)]
        return merge_sorted(parts) if merge else parts
    dist.gather(t, None, dst=0)
    return None
