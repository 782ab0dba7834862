#!/bin/bash
# Round profile on the GPU box: plain bench, ncu launch list of the same command, one ncu --set full capture of the streaming kernel
# on the bench-shaped batch (592 config-3 samples).  usage: tools/gpu_profile_round.sh <tag>
tag=$1; mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-extra > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || { tail -5 gpurun_out/${tag}_bench.err; exit 1; }
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/${tag}_ncu_launches.log 2>&1
PROF_UNIQUE=8 PROF_REPS=74 PROF_LAUNCHES=3 python tools/prof_fused.py > gpurun_out/${tag}_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stream4_kernel -s 1 -c 1 -o gpurun_out/${tag}_full -f env PROF_UNIQUE=8 PROF_REPS=74 PROF_LAUNCHES=3 python tools/prof_fused.py > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv
ncu -i gpurun_out/${tag}_full.ncu-rep --page source --csv > gpurun_out/${tag}_full_src.csv
ncu -i gpurun_out/${tag}_full.ncu-rep --page details > gpurun_out/${tag}_full_details.txt
tail -2 gpurun_out/${tag}_ncu_full.log; python tools/show_bench.py gpurun_out/${tag}_bench.json
