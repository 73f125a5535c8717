"""TEST INFRASTRUCTURE ONLY: CPU oracle for the PDLP hot path (see pdlp_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (highs_b200) never does.
"""
