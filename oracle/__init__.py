"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the MIND hot path (ScenePredNet forward + AIME tree step)
used as the parity checker.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package.
The product (mind_b200/) never imports it and has no CPU fallback.

Parity pinning: the restatement is checked against the reference's own
PyTorch modules imported from /root/reference (oracle/ref_loader.py) in
tests/test_oracle_vs_reference.py (runs only where /root/reference exists)
and against golden vectors produced by those modules, committed under
tests/golden/ with the generating script oracle/make_golden.py.
"""
