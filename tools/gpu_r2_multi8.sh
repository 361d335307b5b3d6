#!/bin/bash
# N GPUs of one box (N = first argument): the driver's torchrun launch of bench.py (both arms)
N=${1:-8}; O=gpurun_out/m$N; mkdir -p $O
nvidia-smi -L > $O/smi.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 3 ) > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
tail -c 300 $O/bench_${N}gpu.err
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 5 --warmup 1 ) > $O/bench_${N}gpu_ref.json 2> $O/bench_${N}gpu_ref.err
python - <<PY
import json
d=json.loads(open('$O/bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','gpu_launches')}, d['e2e']['value'], d['latency']['p50'], d['latency']['p99'])
for k,v in d['configs'].items(): print(k, v.get('scaling'), round(v['ms_per_step'],3), round(v['value']), round(v['e2e']['value']), v.get('streams_per_gpu'), v['latency_ms']['p50'], v['latency_ms']['p99'], v['token_check']['identical'])
r=json.loads(open('$O/bench_${N}gpu_ref.json').read().strip().splitlines()[-1]); print('reference', r['value'], r['cpu_baseline']['cores'])
PY
