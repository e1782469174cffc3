#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
