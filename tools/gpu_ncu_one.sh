#!/bin/bash
# ncu --set full of kernels matching $1 (regex), $2 launches -> gpurun_out/$3_summary.txt + $3_raw.csv
# (the .ncu-rep itself, ~1.5 MB per launch, is kept only with KEEP_REP=1: gpurun returns at most 64 MiB)
K=$1; C=${2:-1}; O=${3:-prof}
mkdir -p gpurun_out
python tools/ncu_step.py 1 > gpurun_out/${O}_plain.log 2>&1 &&
{ SRC=""; [ "${KEEP_REP:-0}" = "1" ] && SRC="--import-source on"
ncu --set full --clock-control none $SRC --profile-from-start off -k regex:"$K" -c $C -o gpurun_out/$O python tools/ncu_step.py 1 > gpurun_out/${O}_ncu.log 2>&1
}
echo "ncu rc=$?"; tail -1 gpurun_out/${O}_plain.log
python tools/ncu_summary.py gpurun_out/$O.ncu-rep > gpurun_out/${O}_summary.txt 2>&1
ncu -i gpurun_out/$O.ncu-rep --page raw --csv > gpurun_out/${O}_raw.csv 2>/dev/null
[ "${KEEP_REP:-0}" = "1" ] || rm -f gpurun_out/$O.ncu-rep
du -sh gpurun_out
