#!/usr/bin/env python
"""bench.py -- pairwise SNP distance + transmission likelihood sweep on B200 (driver contract).

Metric: site-pair comparisons/s = P*L / t  (P = N(N-1)/2 pairs, L = nominal alignment length), the
metric BASELINE.json names. Workload at every GPU count: BASELINE.json configs[1] -- 10,000
synthetic E. coli-like sequences x 5 Mb, ~1% variable sites, SNP threshold 20, TransCluster
likelihood with sampling dates (SURVEY 8d "C2").

  step          one whole pass of the hot path: ASCII alignment (resident in HBM) -> column masks +
                N planes -> variable-site bit-planes -> all-pairs tile sweep with fused threshold ->
                ordered edge list -> compared-site counts -> transmission likelihood -> edges on host.
  value         P*L / step time, inputs resident in HBM (device-generated synthetic ASCII).
  e2e           same metric through the C-ABI call that takes a HOST buffer (tracs_pairsnp_host):
                H2D copy of the 50 GB ASCII matrix and D2H of the edge list inside the timed region.
  --gpus N      weak scaling over independent objects: one C2-shaped MSA (one reference genome's
                alignment, distinct seed) per GPU -- the loop `for msa in args.msa_files` of
                tracs/distance.py:159 -- no traffic during the sweep, per-MSA edge lists gathered to
                rank 0 over NCCL inside the timed step. (--shard tiles: strong scaling of ONE alignment
                by triangle row-blocks, ingest replicated.)
  --impl reference   the unmodified reference (oracle/_ref, built from /root/reference/src) timed on
                the host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WORKLOAD = dict(name="C2: 10000 seqs x 5 Mb E. coli-like, 1% variable sites, dist<=20, transcluster with dates",
                n=10000, L=5_000_000, p_var=0.01, n_clusters=100, mu=5.0, p_N=1e-3, p_amb=0.0, gc=0.508, seed=2,
                n_days=180, gaps=2, dist=20, lamb=29.903, beta=73.0, threshold_Ek=0.01)
# BASELINE.json configs[2]: only reachable with the site-sharded multi-GPU mode (200 GB of ASCII)
WORKLOAD_C3 = dict(name="C3: 100000 seqs x 2 Mb sparse alignment, 1% variable sites, dist<=20, site-sharded over the GPUs",
                   n=100000, L=2_000_000, p_var=0.01, n_clusters=2000, mu=5.0, p_N=1e-3, p_amb=0.0, gc=0.5, seed=3,
                   n_days=180, gaps=2, dist=20, lamb=29.903, beta=73.0, threshold_Ek=0.01)
SECONDS_IN_YEAR = 31556952.0


def n_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on a bounded sample
# ------------------------------------------------------------------------------------------------
class RefSample:
    """Bounded sample of the workload for the CPU reference: same generator, fewer/shorter sequences.
    pair stage: n_p x L_p FASTA; the reference's serial loader is timed on the same file through the
    empty-second-FASTA trick (pair loop empty: src/pairsnp.hpp:352-360,395)."""

    def __init__(self, n_p=1024, L_p=100_000):
        from tracs_b200 import synth
        self.n_p, self.L_p = n_p, L_p
        self.dir = tempfile.mkdtemp(prefix="tracs_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        w = WORKLOAD
        seqs = synth.generate(n_p, L_p, p_var=w["p_var"], n_clusters=max(2, w["n_clusters"] * n_p // w["n"]), mu=w["mu"],
                              p_N=w["p_N"], gc=w["gc"], seed=w["seed"], gaps=w["gaps"])
        self.msa = os.path.join(self.dir, "sample.fasta")
        synth.write_fasta(self.msa, seqs)
        self.empty = os.path.join(self.dir, "empty.fasta")
        open(self.empty, "w").close()
        self.desc = ("reference pairsnp (oracle/_ref, unmodified src/pairsnp.hpp, -O3 -ffast-math) on %d seqs x %d bp of the "
                     "same generator; pair stage = full call minus load-only call; value = projected whole-job rate at "
                     "%d x %d: P*L / (N*L/load_rate + P*L/pair_rate)" % (n_p, L_p, w["n"], w["L"]))

    def cleanup(self):
        import shutil
        shutil.rmtree(self.dir, ignore_errors=True)

    def step(self, mod, threads):
        w = WORKLOAD
        t0 = time.perf_counter()
        mod.pairsnp(fasta=[self.msa, self.empty], n_threads=threads, dist=w["dist"], filter=False)
        t_load = time.perf_counter() - t0
        t0 = time.perf_counter()
        r = mod.pairsnp(fasta=[self.msa], n_threads=threads, dist=w["dist"], filter=False)
        t_full = time.perf_counter() - t0
        t_pairs = max(t_full - t_load, 1e-6)
        P_s = self.n_p * (self.n_p - 1) // 2
        pair_rate = P_s * self.L_p / t_pairs
        load_rate = self.n_p * self.L_p / t_load
        P = w["n"] * (w["n"] - 1) // 2
        t_proj = w["n"] * w["L"] / load_rate + P * w["L"] / pair_rate
        return dict(t_step=t_load + t_full, pair_rate=pair_rate, load_rate=load_rate, projected=P * w["L"] / t_proj,
                    edges=len(r[0]))


def load_reference():
    from oracle import refmod
    if refmod.available():
        mod, variant = refmod.load()
        return mod, "reference", "oracle/_ref/%s" % variant
    # the oracle port (plain C restatement) when the reference could not be compiled
    from oracle import oracle

    class _Port:
        @staticmethod
        def pairsnp(fasta, n_threads, dist, filter):
            return oracle.pairsnp(fasta, n_threads=n_threads, dist=dist, filter=filter, as_lists=False)
    return _Port, "port", "oracle/liboracle.so"


def run_reference(args, saved_stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    mod, kind, what = load_reference()
    cores = n_cores()
    smp = RefSample()
    try:
        for _ in range(args.warmup):
            smp.step(mod, cores)
        t0 = time.perf_counter()
        rs = [smp.step(mod, cores) for _ in range(args.steps)]
        t = time.perf_counter() - t0
    finally:
        smp.cleanup()
    val = float(np.median([r["projected"] for r in rs]))
    line = {
        "impl": "reference", "metric": "site-pair comparisons/s (P*L/t)", "value": val, "unit": "site-pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 bitset words", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "sample": smp.desc},
        "cpu_baseline": {"value": val, "unit": "site-pairs/s", "cores": cores, "kind": kind, "sample": smp.desc, "what": what,
                         "pair_stage_site_pairs_per_s": float(np.median([r["pair_rate"] for r in rs])),
                         "load_bases_per_s": float(np.median([r["load_rate"] for r in rs]))},
        "e2e": {"value": val, "unit": "site-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(saved_stdout, line)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock / throttle-reason sampler for the timed region. In-process NVML (nvidia_ml_py) polled from a
    thread: spawning `nvidia-smi -lms` inside a 250 ms timed region costs more than the region itself
    (its start-up holds driver locks and showed up as 50-200 ms stalls of single steps). The sampler is
    initialised before warm-up; only samples taken between mark() and stop() are reported."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period_s=0.02):
        self.index, self.period = index, period_s
        self.rows, self.t_mark = [], None
        self.ok, self.stop_flag = False, False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as ex:
            self.err = repr(ex)

    def start(self):
        if not self.ok:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def mark(self):
        self.t_mark = time.perf_counter()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag = True
        self.t.join(timeout=2)
        rows = [r for r in self.rows if self.t_mark is None or r[0] >= self.t_mark]
        if not rows:
            rows = self.rows[-1:]
        reasons = sorted({nm for nm, bit in self.REASONS.items() for r in rows if r[2] & bit})
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(rows), "power_w_max": max([r[3] for r in rows]) if rows else None, "source": "nvml, 20 ms period"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def _claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries (NCCL's version banner, for one) print to fd 1,
    so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _emit(saved_fd, line):
    sys.stdout.flush()
    os.write(saved_fd, (json.dumps(line) + "\n").encode())


def run_sites(args, saved_stdout, torch, tracs_b200, dist_mod, device, rank, world, local):
    """Strong scaling of ONE alignment too large for a single GPU (BASELINE configs[2]): rank r holds the
    column slab r of every sequence; see tracs_b200/sites.py for the exchange steps."""
    from tracs_b200 import sites
    w = dict(WORKLOAD_C3)
    if args.n:
        w["n"] = args.n
        w["n_clusters"] = max(2, args.n // 50)
    if args.L:
        w["L"] = args.L
    n, L = w["n"], w["L"]
    P = n * (n - 1) // 2
    lo, hi = sites.slab_bounds(L, rank, world)
    Ls = hi - lo
    pitch = max(128, (Ls + 127) // 128 * 128)
    slab = torch.empty(n * pitch, dtype=torch.uint8, device=device)
    d_days = torch.empty(n, dtype=torch.int32, device=device)
    tracs_b200.synth_device(slab.data_ptr(), n, Ls, pitch, seed=w["seed"], p_var=w["p_var"], n_clusters=w["n_clusters"], mu=w["mu"],
                            p_N=w["p_N"], p_amb=w["p_amb"], gc=w["gc"], n_days=w["n_days"], gaps=w["gaps"], dev_days=d_days.data_ptr(),
                            site_offset=lo, L_total=L)
    days = d_days.cpu().numpy()

    def step():
        return sites.sweep(torch, dist_mod, device, rank, world, slab.data_ptr(), n, Ls, pitch, L, w["dist"], days=days,
                           lamb=w["lamb"], beta=w["beta"], threshold_Ek=w["threshold_Ek"])

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist_mod.barrier()
            torch.cuda.synchronize()

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step()
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    clocks.mark()
    ev0.record()
    for _ in range(args.steps):
        res, st = step()
        stats.append(st)
    ev1.record()
    sync()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist_mod.all_reduce(t_ms, op=dist_mod.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    if rank == 0:
        def avg(k):
            return float(np.mean([s_[k] for s_ in stats]))
        peak = tracs_b200.int_peak()
        peak_wp = min(peak["lop3_per_s"] / 4.0, peak["popc_per_s"])
        wp = avg("swept_wordpairs")
        pf_words = int(round(wp / max(1.0, avg("n_pairs"))))
        t_sw = avg("ms_sweep") * 1e-3
        if avg("tc_sweep") > 0.5:
            roof = {"bound": "tensor", "kernel": "k_sweep_tc", "what": "prefilter launch on this rank's row-blocks (first %d local words), " % pf_words +
                    "tcgen05 int8 one-hot GEMM", "achieved": 2 * wp * 32 * 4 / t_sw / 1e12, "peak": 4500.0, "unit": "TOP/s",
                    "frac": 2 * wp * 32 * 4 / t_sw / 1e12 / 4500.0, "traffic": None, "ms_per_launch": avg("ms_sweep"),
                    "peak_source": "NOMINAL dense int8 (4.5 POP/s)", "equivalent_int_pipe_frac": (wp / t_sw) / peak_wp}
        else:
            roof = {"bound": "int_pipe", "kernel": "k_sweep", "what": "prefilter launch on this rank's row-blocks (first %d local words)" % pf_words,
                    "achieved": wp * 6 / t_sw / 1e9, "peak": peak_wp * 6 / 1e9, "unit": "Ginstr/s", "frac": (wp / t_sw) / peak_wp,
                    "traffic": None, "ms_per_launch": avg("ms_sweep"), "peak_source": "measured in this run (tracs_int_peak)"}
        line = {
            "metric": "site-pair comparisons/s (P*L/t)", "value": P * L / (ms_per_step * 1e-3), "unit": "site-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 bit-planes (int32 counts), f64 likelihood",
            "data": "synthetic (device-generated ASCII column slabs, seeded)",
            "config": {"workload": w["name"], "n": n, "L": L, "pairs": P, "edges": int(len(res["rows"])), "dist": w["dist"],
                       "candidates": int(stats[-1]["n_candidates_all"]),
                       "parallelism": "site-sharded: rank r ingests columns [L*r/R, L*(r+1)/R) of every sequence and prefilters its "
                                      "row-blocks; candidates all-gathered, per-slab partial d and |N u N| all-reduced (NCCL)",
                       "l2": "inputs (%.1f GB ASCII per GPU) larger than L2" % (n * pitch / 1e9)},
            "clocks": clk, "gpu_launches": int(sum(s_["kernel_launches"] + 1 for s_ in stats)), "roofline": roof,
            "stages_ms": {k: avg(k) for k in ("ms_pack", "ms_compact", "ms_sweep", "ms_sort", "ms_refine", "ms_total")},
            "e2e": {"value": None, "unit": "site-pairs/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "note": "host-buffer e2e is measured on the N=1 C2 line"},
            "cpu_baseline": {"value": None, "note": "timed at N=1 only"},
        }
        _emit(saved_stdout, line)
    if world > 1:
        dist_mod.barrier()
        dist_mod.destroy_process_group()
    return 0


def main():
    saved_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=None, help="override sample count (debug only; invalidates the headline)")
    ap.add_argument("--L", type=int, default=None, help="override alignment length (debug only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--shard", default="msa", choices=["msa", "tiles", "sites"],
                    help="N>1: one MSA per GPU (weak, default) | row-blocks of one MSA, ingest replicated (strong) | "
                         "column slabs of the C3 alignment per GPU (strong; the only mode that fits 100k x 2 Mb)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, saved_stdout)
    args.warmup = max(args.warmup, 3)

    import torch
    import tracs_b200
    from tracs_b200 import _lib
    from tracs_b200.multi import PipelinedGather

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path")
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().tracs_set_device(local))
    device = torch.device("cuda", local)
    dist_mod = None
    gatherer = None
    if world > 1:
        # keep NCCL's own banner / debug lines off stdout: the contract is ONE JSON line there
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=device)
        gatherer = None  # created once by_tiles is known

    if args.shard == "sites":
        return run_sites(args, saved_stdout, torch, tracs_b200, dist_mod, device, rank, world, local)

    w = dict(WORKLOAD)
    if args.n:
        w["n"] = args.n
        w["n_clusters"] = max(2, args.n // 100)
    if args.L:
        w["L"] = args.L
    n, L = w["n"], w["L"]
    P = n * (n - 1) // 2
    pitch = (L + 127) // 128 * 128

    # ---- synthetic input, generated in device memory ------------------------------------------
    seqs = torch.empty(n * pitch, dtype=torch.uint8, device=device)
    d_days = torch.empty(n, dtype=torch.int32, device=device)
    by_tiles = world > 1 and args.shard == "tiles"
    my_seed = w["seed"] if (world == 1 or by_tiles) else w["seed"] + 1000 * rank
    tracs_b200.synth_device(seqs.data_ptr(), n, L, pitch, seed=my_seed, p_var=w["p_var"], n_clusters=w["n_clusters"], mu=w["mu"],
                            p_N=w["p_N"], p_amb=w["p_amb"], gc=w["gc"], n_days=w["n_days"], gaps=w["gaps"], dev_days=d_days.data_ptr())
    days = d_days.cpu().numpy()
    kw = dict(dist=w["dist"], days=days, lamb=w["lamb"], beta=w["beta"], threshold_Ek=w["threshold_Ek"],
              shard_rank=rank if by_tiles else 0, shard_world=world if by_tiles else 1)

    peak = tracs_b200.int_peak() if rank == 0 else None

    if world > 1:
        gatherer = PipelinedGather(torch, dist_mod, device, rank, world, merge=by_tiles)

    def step():
        """One pass of the hot path. At N > 1 the edge-list gather of this step is queued and overlaps
        the next step's sweep; drain() below completes every gather inside the timed region."""
        res = tracs_b200.pairsnp_device(seqs.data_ptr(), n, L, pitch, copy=False, **kw)
        st = tracs_b200.last_stats()
        if world > 1:
            gatherer.submit(res)
        return res, st, res

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist_mod.barrier()
            torch.cuda.synchronize()

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        res, st, merged = step()   # held like in the timed loop, so the result-buffer cache reaches its steady state
    if world > 1:
        gatherer.drain()
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    clocks.mark()
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        res, st, merged = step()
        stats.append(st)
    if world > 1:
        gathered = gatherer.drain()   # every step's gather has landed on rank 0 before the clock stops
        merged = gathered[-1] if rank == 0 else None
    ev1.record()
    sync()
    t_wall = time.perf_counter() - t_wall0
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist_mod.all_reduce(t_ms, op=dist_mod.ReduceOp.MAX)
    ms = float(t_ms.item())
    ms_per_step = ms / args.steps
    n_msa = 1 if (world == 1 or by_tiles) else world
    value = n_msa * P * L / (ms_per_step * 1e-3)

    def avg(k):
        return float(np.mean([s[k] for s in stats]))

    line = None
    if rank == 0:
        st = stats[-1]
        if isinstance(merged, dict):
            n_edges = len(merged["rows"])
        else:
            n_edges = int(sum(len(p["rows"]) for p in merged))
        launches = int(sum(s["kernel_launches"] for s in stats))
        # ---- rooflines ---------------------------------------------------------------------------
        # k_sweep (INT-pipe bound): algorithmic work = 6 INT instructions per 32-site word-pair
        # (1 AND + 3 AND-OR LOP3 + POPC + ADD) over the words the launch sweeps.
        # k_pack (HBM bound): reads the n*L ASCII bytes once, writes the N bit-plane (n*L/8) + summaries.
        peak_wp = min(peak["lop3_per_s"] / 4.0, peak["popc_per_s"])
        hbm = None
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            hbm, hbm_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"
        traffic = {}
        prof = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof))
            except Exception:
                pass

        def tc_roof(wordpairs, ms, what):
            macs = wordpairs * 32 * 4          # algorithmic: one-hot K = 4 per site; the kernel executes K = 5 (N column)
            return {"bound": "tensor", "kernel": "k_sweep_tc", "what": what, "achieved": 2 * macs / (ms * 1e-3) / 1e12, "peak": 4500.0,
                    "unit": "TOP/s", "frac": 2 * macs / (ms * 1e-3) / 1e12 / 4500.0, "traffic": None, "ms_per_launch": ms,
                    "executed_tops": 2 * macs * 1.25 / (ms * 1e-3) / 1e12,
                    "peak_source": "NOMINAL dense int8 (4.5 POP/s); MEASURED_PEAKS.json has no int8 figure. tools/tc_rate.cu measures "
                                   "64.1 clk per M128 N128 K32 MMA = the full 8192 MAC/clk/SM on this part",
                    "equivalent_int_pipe_frac": (wordpairs / (ms * 1e-3)) / peak_wp}

        def sweep_roof(wordpairs, ms, what):
            return {"bound": "int_pipe", "kernel": "k_sweep", "what": what, "achieved": wordpairs * 6 / (ms * 1e-3) / 1e9,
                    "peak": peak_wp * 6 / 1e9, "unit": "Ginstr/s", "frac": (wordpairs * 6 / (ms * 1e-3)) / (peak_wp * 6),
                    "traffic": traffic.get("k_sweep_full_length_dram_bytes_per_launch") if what.startswith("full") else
                    traffic.get("k_sweep_prefilter_dram_bytes_per_launch"),
                    "achieved_wordpairs_per_s": wordpairs / (ms * 1e-3), "ms_per_launch": ms,
                    "peak_source": "measured in this run (tracs_int_peak: register-resident LOP3 and POPC loops; a word-pair needs "
                                   "4 LOP3 on the 64-lane ALU pipe and 1 POPC on the 16-lane XU pipe => min(lop3/4, popc) word-pairs/s)",
                    "lop3_per_s": peak["lop3_per_s"], "popc_per_s": peak["popc_per_s"],
                    "mix_wordpairs_per_s": peak["mix_wordpairs_per_s"], "mix_imad_wordpairs_per_s": peak["mix_imad_wordpairs_per_s"]}

        npitch_words = max(32, ((L + 31) // 32 + 31) // 32 * 32)
        # the main pack launch: k_pack over all samples, or -- early extraction -- k_pack_x over samples 256.. (the first 256
        # are packed by a separate small k_pack launch that yields the early site list); timed alone by the library
        n_early = int(round(avg("n_early_sites")))
        n_main = n - 256 if n_early else n
        pack_kernel = "k_pack_x" if n_early else "k_pack"
        # algorithmic bytes: per base 1 B read + 1/8 B N-plane write (+ summary byte per 1024); k_pack_x also writes one
        # byte per (sample, early site)
        pack_bytes = n_main * L + n_main * npitch_words * 4 + n_main * (npitch_words // 32) + n_main * n_early
        ms_main = avg("ms_pack_main")
        roof_pack = {"bound": "hbm", "kernel": pack_kernel, "achieved": pack_bytes / (ms_main * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": pack_bytes / (ms_main * 1e-3) / 1e9 / hbm, "traffic": traffic.get(pack_kernel + "_dram_bytes_per_launch"),
                     "ms_per_launch": ms_main, "algorithmic_bytes": pack_bytes, "samples_in_launch": n_main, "early_sites": n_early,
                     "peak_source": hbm_src}
        prefiltered = avg("n_candidates") > 0 or avg("ms_refine") > 0
        pf_words = int(round(avg("swept_wordpairs") / max(1.0, avg("n_pairs"))))  # window of the prefilter launch(es), in 32-site words
        on_tc = avg("tc_sweep") > 0.5
        roof_sweep = (tc_roof if on_tc else sweep_roof)(avg("swept_wordpairs"), avg("ms_sweep"),
                                                        "prefilter launch (first %d words of every pair)" % pf_words if prefiltered else "full-length sweep")
        # the same tile kernel forced over the full length (what an unthresholded / dense run executes)
        t_full = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res_full = tracs_b200.pairsnp_device(seqs.data_ptr(), n, L, pitch, full_sweep=True, copy=False, **kw)
            torch.cuda.synchronize()
            t_full.append(time.perf_counter() - t0)
        st_full = tracs_b200.last_stats()
        roof_full = sweep_roof(st_full["swept_wordpairs"], st_full["ms_sweep"], "full-length sweep (prefilter disabled)")
        roof_full["whole_step_ms"] = 1e3 * min(t_full)
        roof_full["whole_step_value"] = n_msa * P * L / min(t_full)
        roof_full["edges_equal_default_path"] = bool(np.array_equal(res_full["rows"], res["rows"]) and np.array_equal(res_full["cols"], res["cols"])
                                                     and np.array_equal(res_full["dist"], res["dist"]))
        # the same full-length sweep on the tensor cores (tcgen05 kind::i8 one-hot GEMM, K = 5 int8 per site)
        roof_tc = None
        try:
            t_tc = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res_tc = tracs_b200.pairsnp_device(seqs.data_ptr(), n, L, pitch, full_sweep="tc", copy=False, **kw)
                torch.cuda.synchronize()
                t_tc.append(time.perf_counter() - t0)
            st_tc = tracs_b200.last_stats()
            roof_tc = tc_roof(st_tc["swept_wordpairs"], st_tc["ms_sweep"], "full-length sweep, tcgen05.mma kind::i8 M128 N128 K32, operands "
                              "expanded from the bit-planes in shared memory, int32 accumulators in TMEM")
            roof_tc.update({"whole_step_ms": 1e3 * min(t_tc), "speedup_vs_int_pipe_kernel": st_full["ms_sweep"] / st_tc["ms_sweep"],
                       "edges_equal_default_path": bool(np.array_equal(res_tc["rows"], res["rows"]) and np.array_equal(res_tc["cols"], res["cols"])
                                                        and np.array_equal(res_tc["dist"], res["dist"]))})
        except Exception as ex:
            roof_tc = {"kernel": "k_sweep_tc", "error": repr(ex)}
        roof = roof_pack if avg("ms_pack") >= avg("ms_sweep") else roof_sweep
        stages = {k: avg(k) for k in ("ms_pack", "ms_pack_main", "ms_compact", "ms_sweep", "ms_refine", "ms_sort", "ms_ncomp", "ms_trans", "ms_d2h", "ms_total")}
        stages["n_candidates"] = avg("n_candidates")
        stages["per_step_ms_total"] = [round(s_["ms_total"], 2) for s_ in stats]
        stages["per_step_ms_d2h"] = [round(s_["ms_d2h"], 2) for s_ in stats]
        line = {
            "metric": "site-pair comparisons/s (P*L/t)", "value": value, "unit": "site-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if by_tiles else "weak", "vs_baseline": None,
            "dtype": "u32 bit-planes (int32 counts), f64 likelihood", "data": "synthetic (device-generated ASCII alignment, seeded)",
            "config": {"workload": w["name"], "n": n, "L": L, "pairs": P, "variable_sites": int(st["n_variable_sites"]),
                       "words": int(st["n_words"]), "edges": int(n_edges), "dist": w["dist"],
                       "msas": n_msa,
                       "parallelism": ("triangle row-blocks of one MSA dealt boustrophedon over %d GPUs; ingest replicated" % world) if by_tiles
                       else ("%d independent MSA(s), one per GPU (tracs/distance.py:159 loop); edge lists gathered to rank 0" % world),
                       "algorithm": "exact filter-and-refine: tile sweep over the first %d words of every pair, per-pair refinement of the " % pf_words +
                                    "survivors; roofline_kernels.k_sweep_full_length gives the same step with the full-length tile sweep",
                       "l2": "inputs (%.1f GB ASCII) larger than L2; no flush needed" % (n * pitch / 1e9)},
            "clocks": clk, "gpu_launches": launches, "roofline": roof,
            "roofline_kernels": {pack_kernel: roof_pack, "k_sweep": roof_sweep, "k_sweep_full_length": roof_full, "k_sweep_tc_full_length": roof_tc},
            "stages_ms": stages,
            "wall_ms_per_step": 1e3 * t_wall / args.steps,
        }

    # ---- e2e: host buffer through the C ABI (N=1 only: the path has one host) ----------------------
    if rank == 0 and world == 1 and not args.no_e2e:
        e2e = {"value": None, "unit": "site-pairs/s", "h2d_bytes_per_step": n * L, "d2h_bytes_per_step": None}
        try:
            host = None
            try:
                host = torch.empty((n, L), dtype=torch.uint8, pin_memory=True)
                e2e["host_memory"] = "pinned"
            except Exception:
                host = torch.empty((n, L), dtype=torch.uint8)
                e2e["host_memory"] = "pageable"
            # fill the host buffer with the same alignment (setup, untimed)
            rows = 500
            for r0 in range(0, n, rows):
                r1 = min(n, r0 + rows)
                host[r0:r1].copy_(seqs[r0 * pitch:r1 * pitch].view(r1 - r0, pitch)[:, :L])
            torch.cuda.synchronize()
            del seqs
            torch.cuda.empty_cache()
            hp = host.numpy()
            o, keep = tracs_b200.api.make_opts(**kw)

            def e2e_step():
                e = _lib.Edges()
                _lib.check(_lib.lib().tracs_pairsnp_host(hp.ctypes.data, n, L, L, C.byref(o), C.byref(e)))
                return _lib.take_edges(e, names=False, copy=False)

            e2e_step()
            torch.cuda.synchronize()
            ev0.record()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                r2 = e2e_step()
            ev1.record()
            torch.cuda.synchronize()
            te = (time.perf_counter() - t0) / args.e2e_steps
            st2 = tracs_b200.last_stats()
            e2e["value"] = P * L / te
            e2e["ms_per_step"] = te * 1e3
            e2e["steps"] = args.e2e_steps
            e2e["d2h_bytes_per_step"] = int(st2["d2h_bytes"])
            e2e["edges_equal_device_path"] = bool(np.array_equal(r2["rows"], res["rows"]) and np.array_equal(r2["dist"], res["dist"])
                                                  and np.array_equal(r2["ncomp"], res["ncomp"]))
            e2e["api"] = "tracs_pairsnp_host (C ABI, host ASCII matrix) incl. fused transmission likelihood"
            del host
        except Exception as ex:  # report, never fake
            e2e["error"] = repr(ex)
        line["e2e"] = e2e
    elif rank == 0:
        line["e2e"] = {"value": None, "unit": "site-pairs/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                       "note": "measured at N=1 only (one host buffer)" if world > 1 else "skipped (--no-e2e)"}

    # ---- cpu baseline beside it (rank 0, N=1) -----------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            mod, kind, what = load_reference()
            smp = RefSample()
            try:
                r = smp.step(mod, n_cores())
            finally:
                smp.cleanup()
            line["cpu_baseline"] = {"value": r["projected"], "unit": "site-pairs/s", "cores": n_cores(), "kind": kind, "sample": smp.desc,
                                    "what": what, "pair_stage_site_pairs_per_s": r["pair_rate"], "load_bases_per_s": r["load_rate"]}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "error": repr(ex)}
    elif rank == 0:
        line["cpu_baseline"] = {"value": None, "note": "timed at N=1 only"}

    if rank == 0:
        _emit(saved_stdout, line)
    if world > 1:
        gatherer.close()
        dist_mod.barrier()
        dist_mod.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
