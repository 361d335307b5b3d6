#!/bin/bash
# One GPU call: parity tests, bench, ncu launch list + full capture of the top kernels. Outputs under gpurun_out/<tag>/
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; RC=$?; echo "pytest rc=$RC" | tee -a $OUT/pytest.log; tail -3 $OUT/pytest.log
if [ $RC -ne 0 ]; then NSB_NO_PDL=1 python -m pytest tests -m gpu -x -q > $OUT/pytest_nopdl.log 2>&1; echo "pytest (no PDL) rc=$?"; tail -3 $OUT/pytest_nopdl.log; fi
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
if [ "${NCU:-1}" != "0" ]; then
python tools/ncu_step.py 2 > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches.csv python tools/ncu_step.py 2 > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
fi
if [ "${NCU:-1}" = "2" ]; then
python tools/ncu_step.py 1 > $OUT/plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"${NCU_K:-gemm_tc|attention|conv_module|layernorm}" -c ${NCU_C:-16} -o $OUT/prof_layer python tools/ncu_step.py 1 > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"rnnt_decode" -c 1 -o $OUT/prof_decode python tools/ncu_step.py 1 > $OUT/ncu_full2.log 2>&1
echo "ncu decode rc=$?"
fi
