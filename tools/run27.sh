#!/bin/bash
timeout 300 python tools/bwd_diag.py cifar10 128 128 100 2>&1 | tail -30
echo ---- no pair
LSNF_NO_PAIR=1 timeout 300 python tools/bwd_diag.py cifar10 128 128 100 2>&1 | tail -14
echo ---- no tma store, no streamk
LSNF_NO_TMA_STORE=1 LSNF_STREAMK=0 timeout 300 python tools/bwd_diag.py cifar10 128 128 100 2>&1 | tail -14
