"""CPU oracle for the TamaGo self-play hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference arm may import this package.  The product (tamago_b200/) never does.
"""
