#!/bin/bash
export PYTHONUNBUFFERED=1
echo "== default"; python tools/check_train_determinism.py 2>&1 | grep -v Warn | tail -9
echo "== SEG3D_CIN1_TOEPLITZ=0"; SEG3D_CIN1_TOEPLITZ=0 python tools/check_train_determinism.py 2>&1 | grep -v Warn | tail -9
