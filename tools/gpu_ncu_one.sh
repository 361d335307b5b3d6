#!/bin/bash
# ncu --set full of kernels matching $1 (regex), $2 launches, output gpurun_out/$3.ncu-rep
K=$1; C=${2:-1}; O=${3:-prof}
mkdir -p gpurun_out
python tools/ncu_step.py 1 > gpurun_out/${O}_plain.log 2>&1 &&
{ SRC="--import-source on"; [ "${NCU_SRC:-1}" = "0" ] && SRC=""
ncu --set full --clock-control none $SRC --profile-from-start off -k regex:"$K" -c $C -o gpurun_out/$O python tools/ncu_step.py 1 > gpurun_out/${O}_ncu.log 2>&1
}
echo "ncu rc=$?"; tail -1 gpurun_out/${O}_plain.log; ls -la gpurun_out/
