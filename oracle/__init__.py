"""Test-infrastructure oracles (CPU restatements of the reference hot path).  Not product code: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package."""
