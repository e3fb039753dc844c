#!/usr/bin/env python3
"""Per-step kernel shares from an ncu launch list (CSV with gpu__time_duration.sum, optionally dram__bytes_*).

usage: tools/launch_summary.py launches.csv
A step starts at the small first-chunk pack launch (tracs::k_pack for ASCII input, tracs::k_pack4<0, ...> for packed
input); the LAST complete step of the default (filter-and-refine) path is summarised -- bench.py also runs forced
full-length sweeps for roofline_kernels, which are not part of the timed step."""
import csv
import sys
from collections import OrderedDict


def load(path):
    rows = list(csv.reader(open(path, newline="")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ix = {k: i for i, k in enumerate(rows[h])}
    launches = OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= ix["Metric Value"] or not r[0].isdigit():
            continue
        d = launches.setdefault(int(r[0]), {"name": r[ix["Kernel Name"]].split("(")[0].replace("void ", "")})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r[ix["Metric Name"]]] = v * scale
    return list(launches.values())


def is_step_start(name):
    return (name == "tracs::k_pack" or name.startswith("tracs::k_pack4<0>") or name.startswith("tracs::k_pack4<0,")
            or name.startswith("tracs::k_pack4<(bool)0"))


def main():
    L = load(sys.argv[1])
    names = [d["name"] for d in L]
    starts = [i for i, k in enumerate(names) if is_step_start(k)]
    if len(starts) < 2:
        raise SystemExit("fewer than two first-chunk pack launches in the list")
    spans = [(starts[i], starts[i + 1]) for i in range(len(starts) - 1)]
    default = [sp for sp in spans if any(k in ("tracs::k_refine", "tracs::k_pairs_gather") for k in names[sp[0]:sp[1]])]
    lo, hi = (default or spans)[-1]
    agg = OrderedDict()
    for d in L[lo:hi]:
        c, s, b = agg.get(d["name"], (0, 0.0, 0.0))
        agg[d["name"]] = (c + 1, s + d.get("gpu__time_duration.sum", 0.0), b + d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0))
    tot = sum(s for _, s, _ in agg.values())
    have_bytes = any(b for _, _, b in agg.values())
    print(f"One step = {hi - lo} launches, {tot / 1e6:.3f} ms of kernel time\n")
    print("| kernel | launches / step | ms / step | share |" + (" DRAM GB / step |" if have_bytes else ""))
    print("|---|---|---|---|" + ("---|" if have_bytes else ""))
    for k, (c, s, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[:90]} | {c} | {s / 1e6:.3f} | {s / tot:.3f} |" + (f" {b / 1e9:.2f} |" if have_bytes else ""))


if __name__ == "__main__":
    main()
