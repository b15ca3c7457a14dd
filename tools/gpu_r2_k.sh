#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=8 2>&1 | tail -40 ) 2>&1 | tail -50
