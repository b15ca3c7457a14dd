#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_shim.py tests/test_parity_measure.py tests/test_bindings_ops.py tests/test_sparse_vjp.py -m gpu -q 2>&1 | tail -12
timeout 300 python tools/e2e_probe.py 30 2>&1 | tail -3
