#!/bin/bash
echo "--- FLUSH=2"; DIAG_BIG=1 timeout 300 python scripts/diag_in.py 2>&1 | tail -14
echo "--- FLUSH=8"; DIAG_BIG=1 GNNFD_B200_LIB=$PWD/gnn_fraud_detection_b200/libgnnfd_b200_f8.so timeout 300 python scripts/diag_in.py 2>&1 | tail -14
