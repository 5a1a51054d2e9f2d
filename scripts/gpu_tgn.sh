#!/bin/bash
mkdir -p gpurun_out
G=${1:-1}
if [ $G = 1 ]; then
timeout 600 python bench.py --workload tgn_snapshots --steps 20 --warmup 3 > gpurun_out/bench_tgn_1.json 2> gpurun_out/bench_tgn_1.err; echo "tgn 1 exit $?"; tail -2 gpurun_out/bench_tgn_1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tgn_1.json')); print(d['ms_per_step'], d['value'], d['timing'], d['cpu_baseline'], d['gpu_launches'])"
else
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $G --workload tgn_snapshots --steps 20 --warmup 3 > gpurun_out/bench_tgn_$G.json 2> gpurun_out/bench_tgn_$G.err; echo "tgn $G exit $?"; tail -3 gpurun_out/bench_tgn_$G.err
tail -1 gpurun_out/bench_tgn_$G.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['timing'], d['config']['local_nodes'])"
fi
