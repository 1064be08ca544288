"""CPU oracle for RSLO's per-frame-pair hot path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package ``rslo_b200`` never imports it.
"""
