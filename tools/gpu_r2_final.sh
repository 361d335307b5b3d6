#!/bin/bash
# round 2, final evidence run: full GPU suite, smoke, bench (+ reference arm), ncu (launch list, --set full of the hot kernels)
O=gpurun_out/fin; mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
( time timeout 900 python bench.py ) > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.err
( time timeout 600 python bench.py --impl reference ) > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
# ncu launch list, config 2 (single steps: full-width decode) and config 3
timeout 200 python tools/ncu_step.py 2 > $O/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_cfg2.csv python tools/ncu_step.py 2 > $O/ncu_l2.log 2>&1
echo "launch list cfg2 rc=$?"
NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_cfg3.csv python tools/ncu_step.py 1 > $O/ncu_l3.log 2>&1
echo "launch list cfg3 rc=$?"
# ncu --set full
full() { # regex, count, tag, env...
  local K=$1 C=$2 T=$3; shift 3
  env "$@" timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"$K" -c $C -o $O/$T python tools/ncu_step.py 1 > $O/${T}_ncu.log 2>&1
  echo "ncu $T rc=$?"
  python tools/ncu_summary.py $O/$T.ncu-rep > $O/${T}_summary.txt 2>&1
  ncu -i $O/$T.ncu-rep --page raw --csv > $O/${T}_raw.csv 2>/dev/null
  rm -f $O/$T.ncu-rep
}
full "gemm_tc" 10 full_gemm_cfg2 X=1
full "attention|conv_module|layernorm|rnnt_decode|logmel|stem_conv0" 8 full_misc_cfg2 X=1
full "gemm_tc_pair256" 8 full_gemm_cfg3 NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16
full "attention|conv_module|layernorm_rows|rnnt_decode|dequant" 8 full_misc_cfg3 NSB_BENCH_STREAMS=256 NSB_BENCH_R=6 NSB_BENCH_COMPUTE=q8_0 NSB_BENCH_KV=f16
timeout 300 python tools/trace_step.py 2 > $O/trace_cfg2.txt 2>&1; grep -A16 "per kernel class" $O/trace_cfg2.txt | cut -c1-170
du -sh $O
