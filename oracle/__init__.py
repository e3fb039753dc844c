"""CPU oracle for the TRACS pairwise-distance hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product (tracs_b200) never does and has no CPU fallback."""
