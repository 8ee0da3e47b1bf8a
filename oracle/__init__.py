"""CPU oracle for the read-anchoring pass -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (anchored_fusion_b200/) never does.
"""
