"""CPU checks of the small helper scripts that produce committed evidence."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_summary_reads_the_committed_launch_list():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"),
                          os.path.join(ROOT, "profiles", "r1_launches.csv")], capture_output=True, text=True, check=True).stdout
    assert out.startswith("One step = ")
    rows = [ln for ln in out.splitlines() if ln.startswith("| tracs::")]
    names = [ln.split("|")[1].strip() for ln in rows]
    # the dominant kernels of the default (filter-and-refine) step, largest first
    assert names[0] in ("tracs::k_pack_x", "tracs::k_pack") and "tracs::k_refine" in names and "tracs::k_ncomp" in names
    shares = [float(ln.split("|")[4]) for ln in rows]
    assert shares == sorted(shares, reverse=True) and 0.99 < sum(float(ln.split("|")[4]) for ln in out.splitlines()
                                                                 if ln.startswith("| ") and not ln.startswith("| kernel")) < 1.01
