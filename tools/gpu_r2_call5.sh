#!/bin/bash
O=gpurun_out/c5; mkdir -p $O
CF="0:0:-1,112:97,160:97,208:97,256:97,256:2,128:3,128:4,256:97:2,256:97:4,208:97:2,160:97:2,112:97:2"
ROWS=896 R=13 CFGS=$CF timeout 600 python tools/gemm_large.py > $O/gemm_896.txt 2>&1
ROWS=1792 R=6 CFGS=$CF timeout 600 python tools/gemm_large.py > $O/gemm_1792.txt 2>&1
cat $O/gemm_896.txt $O/gemm_1792.txt
