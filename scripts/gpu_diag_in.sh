#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/diag_in.py > gpurun_out/diag_in.log 2>&1; echo "diag exit $?"
tail -60 gpurun_out/diag_in.log
DIAG_QUICK=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/diag_in.py > gpurun_out/diag_in_memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/diag_in_memcheck.log | head -20
