#!/bin/bash
# ncu --set full capture of the large-image kernel (832x832 heads, 10140 cells per image) -> gpurun_out/TAG_*
#   gpurun --timeout 600 -- 'bash profiles/gpu_round_large.sh TAG'
TAG=${1:-large}
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:decode_nms_large -s 3 -c 1 -f -o $O/${TAG}_full_large \
    python profiles/run_profile.py --workload cfg5_832 --iters 5 > $O/${TAG}_ncu_large.log 2>&1
ncu -i $O/${TAG}_full_large.ncu-rep --page raw --csv > $O/${TAG}_raw_large.csv 2>/dev/null
python profiles/ncu_summary.py < $O/${TAG}_raw_large.csv > $O/${TAG}_summary_large.txt 2>&1
grep -E "==|gpu__time_duration|dram__bytes|issue_active.avg.pct|registers_per_thread \[|inst_executed.sum" $O/${TAG}_summary_large.txt | head -8
ncu -i $O/${TAG}_full_large.ncu-rep --page source --print-source cuda,sass --csv > $O/${TAG}_src_large.csv 2>/dev/null
python profiles/ncu_lines.py $O/${TAG}_src_large.csv 1.5 | head -40
