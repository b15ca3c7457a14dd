"""TEST INFRASTRUCTURE — not product code.

oracle/ holds (1) lq_ref.py: a ctypes view of oracle/_ref/liblq_ref.so, the reference's own
lightning.qubit C++ core compiled unmodified from /root/reference (oracle/Makefile, ref_shim.cpp),
and (2) np_oracle.py: a numpy restatement of the same algorithms for small qubit counts.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (pennylane-lightning_b200/) never does.
"""
