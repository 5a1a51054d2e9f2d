#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
echo "--- default run $i"; timeout 300 python scripts/diag_in.py 2>&1 | grep -E "BAD|WORST" | head -8
done
for i in 1 2; do
echo "--- PACK=0 run $i"; GNNFD_BWD_PACK=0 timeout 300 python scripts/diag_in.py 2>&1 | grep -E "BAD|WORST" | head -8
done
for i in 1 2 3; do
echo "--- GD_SYNC run $i"; GNNFD_B200_LIB=$PWD/gnn_fraud_detection_b200/libgnnfd_b200_gdsync.so timeout 300 python scripts/diag_in.py 2>&1 | grep -E "BAD|WORST" | head -8
done
