#!/bin/bash
# nemotron-asr-serve on the GPU box: C++ host over the C ABI, synthetic 24-layer model (speech-rate calibration).
#   leg 1: config 2 shape (64 streams x 160 ms, bf16) at full speed, two steps in flight -> RTFx
#   leg 2: config 4 shape per GPU (128 streams x 80 ms, bf16), real-time pacing -> p50 / p99 chunk latency
OUT=gpurun_out/${1:-serve}; mkdir -p $OUT
M1=$(python -c "import sys; sys.path.insert(0,'tools'); import synth; print(synth.cached_model('f16', 24, R=1, profile='speech'))")
M0=$(python -c "import sys; sys.path.insert(0,'tools'); import synth; print(synth.cached_model('f16', 24, R=0, profile='speech'))")
S=nemotron-speech.cpp_b200/nemotron-asr-serve
timeout 60 $S $M1 --right-context 1 --compute bf16 --max-streams 64 --synthetic 64 30 > $OUT/cfg2.out 2> $OUT/cfg2.err; echo "cfg2 rc=$?"; tail -6 $OUT/cfg2.err
timeout 60 $S $M0 --right-context 0 --compute bf16 --max-streams 128 --realtime --synthetic 128 6 > $OUT/cfg4_rt.out 2> $OUT/cfg4_rt.err; echo "cfg4 rc=$?"; tail -7 $OUT/cfg4_rt.err
