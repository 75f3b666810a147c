"""Checker scripts that need a GPU box and are not collected by pytest: N-rank parity
(dist_check.py, run under torchrun), the randomised parity sweep (fuzz_parity.py) and
the float-noise attribution (noise_check.py).  They live under tests/ because they call
the CPU oracle, which only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may do."""
