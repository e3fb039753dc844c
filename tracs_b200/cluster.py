"""Host-side mirror of the reference's `tracs cluster` stage (tracs/cluster.py:82-139): single-linkage
clusters = connected components of the edges whose chosen column is <= threshold, written as
`sample,cluster`. Names are interned in order of first appearance over ALL rows (tracs/cluster.py:105-108),
labels are numbered like scipy's connected_components; the component search runs on the GPU
(tracs_connected_components)."""
import argparse

import numpy as np

from . import api

COLUMN = {"snp": 3, "filter": 6, "direct": 4, "expectedK": 5}  # tracs/cluster.py:90-97


def cluster(distance_file, output_file, threshold, distance):
    col = COLUMN[distance]
    index, I, J, count = {}, [], [], 0
    with open(distance_file) as f:
        next(f)
        for line in f:
            parts = line.strip().split(",")
            i = index.setdefault(parts[0], len(index))
            j = index.setdefault(parts[1], len(index))
            if float(parts[col]) <= threshold:
                I.append(i)
                J.append(j)
            count += 1
    if count <= 0:
        return None  # the reference logs a warning and writes nothing (tracs/cluster.py:115-117)
    names = list(index.keys())
    n_components, labels = api.connected_components(np.array(I, np.uint64), np.array(J, np.uint64), len(names))
    with open(output_file, "w") as out:
        out.write("sample,cluster\n")
        for nm, lab in zip(names, labels.tolist()):
            out.write(nm + "," + str(lab) + "\n")
    return n_components


def main(argv=None):
    ap = argparse.ArgumentParser(description="Single-linkage transmission clusters (B200)")
    ap.add_argument("-d", "--distances", dest="distance_file", required=True)
    ap.add_argument("-o", "--output", dest="output_file", required=True)
    ap.add_argument("-c", "--threshold", type=float, required=True)
    ap.add_argument("-D", "--distance", choices=list(COLUMN), required=True)
    a = ap.parse_args(argv)
    cluster(a.distance_file, a.output_file, a.threshold, a.distance)


if __name__ == "__main__":
    main()
