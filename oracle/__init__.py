"""CPU oracle for the amcl3d hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (amcl3d_b200) never does.
"""
