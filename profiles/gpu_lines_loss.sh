#!/bin/bash
# per-source-line instruction counts of target_loss_kernel (head 1, N=512, 100 GT)
TAG=${1:-losslines}
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:target_loss_kernel -s 13 -c 1 -f -o $O/${TAG}_full python profiles/loss_time.py 512 100 > $O/${TAG}_ncu.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_raw.csv 2>/dev/null
python profiles/ncu_summary.py < $O/${TAG}_raw.csv > $O/${TAG}_summary.txt 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page source --print-source cuda,sass --csv > $O/${TAG}_src.csv 2>/dev/null
python profiles/ncu_lines.py $O/${TAG}_src.csv 0.7 > $O/${TAG}_lines.txt 2>&1
rm -f $O/${TAG}_full.ncu-rep
grep -E "gpu__time_duration|issue_active.avg.pct|inst_executed.sum " $O/${TAG}_summary.txt | head -5
cat $O/${TAG}_lines.txt | head -70
