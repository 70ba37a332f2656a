#!/bin/bash
# Final evidence run of round 2 (one GPU): ncu launch list of the bench command, --set full summaries + per-function
# instruction counts of the final fused kernel (dense + sparse), then tests, smoke and the two bench arms untouched by ncu.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_final_r02.sh'
O=gpurun_out
mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-extra > $O/r02f_bench_under_ncu.log 2>&1
python profiles/launch_share.py $O/r02f_launches.csv > $O/r02f_launch_list.txt 2>&1
bash profiles/gpu_lines2.sh r02f "cfg2 cfg2_sparse" > /dev/null 2>&1
for w in cfg2 cfg2_sparse; do python profiles/ncu_functions.py $O/r02f_src_$w.csv > $O/r02f_functions_$w.txt 2>&1; done
rm -f $O/r02f_src_*.csv $O/r02f_raw_*.csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 20 --warmup 3 > $O/r02f_bench_reference.json 2>/dev/null
python bench.py --steps 20 --warmup 3 > $O/r02f_bench_n1.json 2> $O/r02f_bench_n1.err
tail -c 300 $O/r02f_bench_n1.err
head -12 $O/r02f_launch_list.txt; head -12 $O/r02f_functions_cfg2.txt; grep -i "issue_active.avg\|gpu__time_duration.sum\|dram__bytes_read.sum \|dram__bytes_write.sum \|inst_executed.sum " $O/r02f_summary_cfg2.txt $O/r02f_summary_cfg2_sparse.txt
