#!/usr/bin/env python3
"""Per-step kernel shares from an ncu launch list (gpu__time_duration.sum CSV).

usage: tools/launch_summary.py launches.csv [steps_in_run]
A step starts at every tracs::k_pack launch; the LAST complete step of the default (filter-and-refine) path,
i.e. one that launches tracs::k_refine, is summarised -- bench.py also runs forced full-length sweeps for
roofline_kernels, which are not part of the timed step."""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 14 and r[0].isdigit()]
    names = [r[4].split("(")[0] for r in rows]
    ns = [float(r[14]) for r in rows]
    starts = [i for i, k in enumerate(names) if k == "tracs::k_pack"]
    if len(starts) < 2:
        raise SystemExit("fewer than two k_pack launches in the list")
    spans = [(starts[i], starts[i + 1]) for i in range(len(starts) - 1)]
    spans = [sp for sp in spans if "tracs::k_refine" in names[sp[0]:sp[1]]] or spans
    lo, hi = spans[-1]
    agg = OrderedDict()
    for k, t in zip(names[lo:hi], ns[lo:hi]):
        c, s = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, s + t)
    tot = sum(s for _, s in agg.values())
    print(f"One step = {hi - lo} launches, {tot / 1e6:.3f} ms of kernel time\n")
    print("| kernel | launches / step | ms / step | share |\n|---|---|---|---|")
    for k, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[:90]} | {c} | {s / 1e6:.3f} | {s / tot:.3f} |")


if __name__ == "__main__":
    main()
