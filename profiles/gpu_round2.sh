#!/bin/bash
# Round-2 evidence run (one GPU): micro-benchmark, ncu launch list of the bench command, per-line / per-function
# instruction counts and --set full summaries of the final fused kernel (dense + sparse), gather-path cost.
#   gpurun --timeout 1200 -- 'bash profiles/gpu_round2.sh'
O=gpurun_out
mkdir -p $O
(cd profiles/micro && ./pair_loop.bin) > $O/r02_micro_pair_loop.txt 2>&1
python profiles/gather_perf.py > $O/r02_gather_perf.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-extra > $O/r02_bench_under_ncu.log 2>&1
python profiles/launch_share.py $O/r02_launches.csv > $O/r02_launch_list.txt 2>&1
bash profiles/gpu_lines2.sh r02v16 "cfg2 cfg2_sparse"
python bench.py --steps 20 --warmup 3 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
python bench.py --steps 200 --warmup 20 --no-extra > $O/r02_bench_n1_200steps.json 2>> $O/r02_bench_n1.err
tail -c 400 $O/r02_bench_n1.err
cat $O/r02_launch_list.txt | head; cat $O/r02_gather_perf.txt
