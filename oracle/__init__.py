"""CPU oracle for the LIA_RAL hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product (lia_ral_b200) never does.
"""
