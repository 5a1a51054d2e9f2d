#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/pytest_gpu_${1:-x}.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_${1:-x}.log
