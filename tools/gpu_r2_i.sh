#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense_mma.py tests/test_sparse_vjp.py tests/test_parity_gates.py tests/test_bindings_ops.py tests/test_conformance.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/dense_bench.py 30 2>&1 | tail -12
