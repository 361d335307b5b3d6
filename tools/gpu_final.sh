#!/bin/bash
# Round-end evidence run: parity suite, smoke, bench (+ reference arm), ncu launch list, in-graph traces of every config, ncu --set full summaries
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-260 $OUT/bench.json
python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference rc=$?"; cut -c1-200 $OUT/bench_reference.json
python tools/ncu_step.py 2 > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches.csv python tools/ncu_step.py 2 > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/trace_step.py 3 > $OUT/trace.txt 2>&1; echo "trace rc=$?"; grep "decode:" $OUT/trace.txt; head -1 $OUT/trace.txt
bash tools/gpu_ncu_one.sh "gemm_tc" 9 $TAG/ncu_gemm
bash tools/gpu_ncu_one.sh "attention|conv_module|layernorm" 5 $TAG/ncu_misc
bash tools/gpu_ncu_one.sh "rnnt_decode|logmel|stem_conv0" 3 $TAG/ncu_decode
bash tools/gpu_traces.sh $TAG | grep -E "^==|^# step|decode:"
bash tools/gpu_configs.sh > $OUT/configs.txt 2>&1; cat $OUT/configs.txt
du -sh gpurun_out
