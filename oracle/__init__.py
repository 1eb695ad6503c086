"""ORACLE — test infrastructure only (see oracle/onnx_interp.py and oracle/host_oracle.c headers).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
PARITY STATUS: parity unpinned — the reference has no tests, golden vectors or runnable backend.
"""
