"""Multi-GPU plumbing for the pair sweep: one process per GPU.

Two ways to shard (both without any traffic during the sweep):
  * one MSA per rank -- the loop `for msa in args.msa_files` of tracs/distance.py:159 (weak scaling);
  * row-blocks of ONE alignment dealt to ranks (tracs_opts_t.shard_rank/shard_world, strong scaling).
Either way the only exchange is ONE variable-length gather of the edge columns to rank 0.
torch.distributed is the transport (NCCL on GPUs; gloo in the CPU tests). Columns travel in a
compact 32 B/edge layout (u32 row, col, d, compared sites; f64 log p0, E[K]) through a cached
page-locked staging buffer."""
import numpy as np

# host-staged path: every column of the edge table (44 B/edge); GPU-to-GPU path: the library's packed block
# (tracs_edges_t.dev_packed, 32 B/edge: rows, cols, RAW dist, ncomp, p0_log, eK -- no filt / datediff, so it is
# only taken when the recombination filter is off; datediff is then recomputed by the caller if needed)
COLUMNS = ("rows", "cols", "dist", "ncomp", "filt", "p0_log", "eK", "datediff")
_DTYPES = (np.uint32, np.uint32, np.uint32, np.uint32, np.uint32, np.float64, np.float64, np.float64)
_BYTES_PER_EDGE = 5 * 4 + 3 * 8
DEV_COLUMNS = ("rows", "cols", "dist", "ncomp", "p0_log", "eK")
_DEV_DTYPES = (np.uint32, np.uint32, np.uint32, np.uint32, np.float64, np.float64)
_DEV_BYTES_PER_EDGE = 4 * 4 + 2 * 8
TILE = 128


def shard_owner(row_block, world):
    """Same deal as the library's shard_owner() (csrc/common.cuh): boustrophedon over ranks."""
    rnd, pos = divmod(row_block, world)
    return (world - 1 - pos) if (rnd & 1) else pos


def _sections(buf, mx, dtypes=_DTYPES):
    """Typed views of the column sections inside a flat uint8 numpy buffer of bytes_per_edge*mx bytes."""
    out, off = [], 0
    for dt in dtypes:
        nb = np.dtype(dt).itemsize * mx
        out.append(buf[off:off + nb].view(dt))
        off += nb
    return out


def merge_by_rowblock(parts, world):
    """parts[r]: dict of columns from rank r, each sorted by (row, col); row-blocks were dealt by
    shard_owner. Returns the columns concatenated in global (row, col) order without sorting."""
    top = 0
    for p in parts:
        if len(p["rows"]):
            top = max(top, int(p["rows"][-1]))
    n_rb = top // TILE + 1
    bounds = [np.searchsorted(p["rows"], np.arange(n_rb + 1, dtype=np.uint64) * TILE) for p in parts]
    out = {}
    for c in parts[0].keys():
        out[c] = np.concatenate([parts[shard_owner(rb, world)][c][bounds[shard_owner(rb, world)][rb]:bounds[shard_owner(rb, world)][rb + 1]]
                                 for rb in range(n_rb)]) if n_rb else parts[0][c][:0]
    return out


class _DevBytes:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface v2)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class EdgeGather:
    """Reusable gather: keeps its pinned / device staging buffers between steps."""

    def __init__(self, torch, dist, device, rank, world):
        self.torch, self.dist, self.device, self.rank, self.world = torch, dist, device, rank, world
        self.pinned = device.type == "cuda"
        self.send = None
        self.recv = None

    def _host(self, nbytes, old):
        if old is not None and old.numel() >= nbytes:
            return old
        return self.torch.empty(int(nbytes * 1.25) + 64, dtype=self.torch.uint8, pin_memory=self.pinned)

    def gather(self, res, merge=True):
        """All ranks call this. Rank 0 returns the merged dict of columns (merge=True: row-block
        shards of one alignment) or the list of per-rank dicts (merge=False: one MSA per rank);
        other ranks return None."""
        torch, dist = self.torch, self.dist
        n_loc = len(res["rows"])
        cnt = torch.tensor([n_loc], dtype=torch.int64, device=self.device)
        cnts = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
        dist.all_gather(cnts, cnt)
        cnts = [int(c.item()) for c in cnts]
        dp = res.get("dev_packed") if hasattr(res, "get") else None
        filt = res.get("filt") if hasattr(res, "get") else None
        has_filt = filt is not None and len(filt) and bool(np.any(np.asarray(filt)))
        if dp is not None and self.device.type == "cuda" and not has_filt:
            return self._gather_device(res, dp, cnts, merge)
        mx = max(cnts + [1])
        mx += mx & 1   # five u32 sections in front of the f64 ones: keep those 8-byte aligned
        nbytes = _BYTES_PER_EDGE * mx
        self.send = self._host(nbytes, self.send)
        secs = _sections(self.send.numpy()[:nbytes], mx)
        for sec, c in zip(secs, COLUMNS):
            v = res.get(c)
            if v is not None and n_loc:
                sec[:n_loc] = v
            elif n_loc:
                sec[:n_loc] = 0
        dsend = self.send[:nbytes].to(self.device, non_blocking=True)
        if self.rank != 0:
            dist.gather(dsend, None, dst=0)
            return None
        bufs = [torch.empty_like(dsend) for _ in range(self.world)]
        dist.gather(dsend, bufs, dst=0)
        self.recv = self._host(nbytes * self.world, self.recv)
        flat = self.recv[:nbytes * self.world]
        for r, b in enumerate(bufs):
            flat[r * nbytes:(r + 1) * nbytes].copy_(b, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        host = flat.numpy()
        parts = []
        for r in range(self.world):
            secs = _sections(host[r * nbytes:(r + 1) * nbytes], mx)
            parts.append({c: s[:cnts[r]] for c, s in zip(COLUMNS, secs)})
        return merge_by_rowblock(parts, self.world) if merge else parts


def _gather_device_impl(self, res, dp, cnts, merge):
    """GPU-to-GPU path: every rank sends its packed device block (tracs_edges_t.dev_packed) straight to
    rank 0 with exact sizes; rank 0 makes ONE device->host copy per rank into its pinned buffer."""
    torch, dist = self.torch, self.dist
    ptr, nbytes = dp
    mine = torch.as_tensor(_DevBytes(ptr, max(nbytes, 1)), device=self.device)[:nbytes] if nbytes else torch.empty(0, dtype=torch.uint8, device=self.device)
    if self.rank != 0:
        if nbytes:
            dist.send(mine, dst=0)
        return None
    bufs = [mine] + [torch.empty(_DEV_BYTES_PER_EDGE * cnts[r], dtype=torch.uint8, device=self.device) for r in range(1, self.world)]
    reqs = [dist.irecv(bufs[r], src=r) for r in range(1, self.world) if cnts[r]]
    for q in reqs:
        q.wait()
    total = _DEV_BYTES_PER_EDGE * sum(cnts)
    self.recv = self._host(total, self.recv)
    off, spans = 0, []
    for r in range(self.world):
        nb = _DEV_BYTES_PER_EDGE * cnts[r]
        if nb:
            self.recv[off:off + nb].copy_(bufs[r], non_blocking=True)
        spans.append((off, nb))
        off += nb
    torch.cuda.current_stream(self.device).synchronize()
    host = self.recv.numpy()
    parts = []
    for r, (o, nb) in enumerate(spans):
        secs = _sections(host[o:o + nb], cnts[r], _DEV_DTYPES)
        parts.append({c: s for c, s in zip(DEV_COLUMNS, secs)})
    return merge_by_rowblock(parts, self.world) if merge else parts


EdgeGather._gather_device = _gather_device_impl


class PipelinedGather:
    """Runs EdgeGather on a worker thread and its own CUDA stream, so the gather of step k overlaps the
    sweep of step k+1 (the library call releases the GIL). submit() per step, drain() before the
    clock stops; results come back in submission order."""

    def __init__(self, torch, dist, device, rank, world, merge):
        import queue
        import threading
        self.g = EdgeGather(torch, dist, device, rank, world)
        self.torch, self.device, self.merge = torch, device, merge
        self.q = queue.Queue()
        self.out = []
        self.err = None
        self.stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        torch = self.torch
        if self.stream is not None:
            torch.cuda.set_device(self.device)
        while True:
            item = self.q.get()
            if item is None:
                self.q.task_done()
                return
            try:
                if self.stream is not None:
                    with torch.cuda.stream(self.stream):
                        self.out.append(self.g.gather(item, merge=self.merge))
                else:
                    self.out.append(self.g.gather(item, merge=self.merge))
            except Exception as ex:  # surfaced by drain()
                self.err = ex
            self.q.task_done()

    def submit(self, res):
        self.q.put(res)

    def drain(self):
        self.q.join()
        if self.err is not None:
            raise self.err
        out, self.out = self.out, []
        return out

    def close(self):
        self.q.put(None)
        self.t.join(timeout=30)


def gather_edges(res, rank, world, dist, torch, device, merge=True):
    """One-shot convenience wrapper around EdgeGather."""
    return EdgeGather(torch, dist, device, rank, world).gather(res, merge=merge)
