"""Drop-in for the reference's compiled extension module `TRACS`
(gtonkinhill/tracs src/python_bindings.cpp:8-26): same four functions, same keyword names,
same return types (Python lists / tuples / ndarray). Put this directory on PYTHONPATH (or call
tracs_b200.install_dropin()) and tracs/distance.py:8, tracs/transcluster.py:2 and
tracs/align.py:21 import it unchanged. All arithmetic runs in libtracs_b200.so (CUDA, sm_100a)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from tracs_b200.api import pairsnp, trans_dist, lprob_k_given_N, calculate_posteriors  # noqa: E402,F401

__doc__ = "Meta Transmission Clustering"
__all__ = ["pairsnp", "trans_dist", "lprob_k_given_N", "calculate_posteriors"]
