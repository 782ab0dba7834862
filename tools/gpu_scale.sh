#!/bin/bash
# usage: tools/gpu_scale.sh <N> <tag>: the bench at N GPUs (own arm, with the extra workloads; config 5 runs at N = 8), then the mini404 strong-scaling line
N=$1; tag=$2; mkdir -p gpurun_out
run() { if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; fi; }
run --steps 20 --warmup 3 > gpurun_out/${tag}_n${N}.json 2> gpurun_out/${tag}_n${N}.err || tail -5 gpurun_out/${tag}_n${N}.err
python tools/show_bench.py gpurun_out/${tag}_n${N}.json
run --steps 20 --warmup 3 --workload mini404 --no-cpu > gpurun_out/${tag}_mini404_n${N}.json 2> gpurun_out/${tag}_mini404_n${N}.err || tail -5 gpurun_out/${tag}_mini404_n${N}.err
tail -1 gpurun_out/${tag}_mini404_n${N}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mini404', {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, d['config'].get('kernel_config'), d['roofline'].get('kernel_ms'))"
