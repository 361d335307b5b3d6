#!/bin/bash
# steady-state rate of the 256-row pair tiles on a GEMM with many waves (7168 rows): is the tile or the layer shape the limit?
O=gpurun_out/c14; mkdir -p $O
CFGS="0:0,256:97,208:97,160:97,256:96,208:96" ROWS=7168 timeout 600 python tools/gemm_large.py > $O/gemm_7168.txt 2>&1
cat $O/gemm_7168.txt
