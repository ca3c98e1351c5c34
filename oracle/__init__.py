"""oracle/ -- TEST INFRASTRUCTURE.  CPU restatements of the reference algorithm for the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; monodetr_b200/ never does (the product path fails loudly without its CUDA library).
"""
