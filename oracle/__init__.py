"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU (plain PyTorch fp32/fp64) restatements of the reference's forward algorithm, the shim +
loader that imports the UNMODIFIED reference from /root/reference when it is present, and the
script that generated the golden vectors under tests/golden/. Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; the product path
(multi-task-transformer_b200/) never does.

Parity pin: the reference ships no tests or golden vectors, so the
restatements are pinned against outputs of the reference itself run in the build container
(tests/test_oracle_vs_reference.py, tests/golden/*.pt made by oracle/make_golden.py).
"""
