#!/bin/bash
# Round-end evidence run: parity suite, smoke, bench (+ reference arm), ncu launch list, in-graph traces of every config,
# ncu --set full summaries (bench config + the large-batch config). Every leg under its own timeout, outputs straight to files.
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-260 $OUT/bench.json
for c in 3 4 5; do timeout 300 python bench.py --config $c --no-cpu-baseline > $OUT/bench_cfg$c.json 2> $OUT/bench_cfg$c.err; echo "bench cfg$c rc=$?"; cut -c1-200 $OUT/bench_cfg$c.json; done
timeout 400 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference rc=$?"; cut -c1-200 $OUT/bench_reference.json
timeout 200 python tools/ncu_step.py 2 > $OUT/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches.csv python tools/ncu_step.py 2 > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 200 python tools/trace_step.py 3 > $OUT/trace.txt 2>&1; echo "trace rc=$?"; grep "decode:" $OUT/trace.txt; head -1 $OUT/trace.txt
timeout 300 bash tools/gpu_ncu_one.sh "gemm_tc" 9 $TAG/ncu_gemm
timeout 300 bash tools/gpu_ncu_one.sh "attention|conv_module|layernorm" 5 $TAG/ncu_misc
timeout 300 bash tools/gpu_ncu_one.sh "rnnt_decode|logmel|stem_conv0" 3 $TAG/ncu_decode
( export NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=f16 NSB_BENCH_KV=f16; timeout 300 bash tools/gpu_ncu_one.sh "gemm_tc|attention_mma" 8 $TAG/ncu_cfg3 )
timeout 600 bash tools/gpu_traces.sh $TAG | grep -E "^==|^# step|decode:"
timeout 600 bash tools/gpu_configs.sh > $OUT/configs.txt 2>&1; cat $OUT/configs.txt
CFGS="0:0,256:2,128:3,128:4,256:97,208:97,160:97,112:97" timeout 200 python tools/gemm_large.py > $OUT/gemm_large_1792.txt 2>&1; tail -8 $OUT/gemm_large_1792.txt
du -sh gpurun_out
