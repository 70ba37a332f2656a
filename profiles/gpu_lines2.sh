#!/bin/bash
# per-source-line instruction counts of the fused kernel (dense + sparse cfg2), no bench run
TAG=${1:-lines}
O=gpurun_out
mkdir -p $O
for w in ${2:-cfg2 cfg2_sparse}; do
  ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 4 -c 1 -f -o $O/${TAG}_full_$w \
      python profiles/run_profile.py --workload $w > $O/${TAG}_ncu_$w.log 2>&1
  ncu -i $O/${TAG}_full_$w.ncu-rep --page raw --csv > $O/${TAG}_raw_$w.csv 2>/dev/null
  python profiles/ncu_summary.py < $O/${TAG}_raw_$w.csv > $O/${TAG}_summary_$w.txt 2>&1
  ncu -i $O/${TAG}_full_$w.ncu-rep --page source --print-source cuda,sass --csv > $O/${TAG}_src_$w.csv 2>/dev/null
  python profiles/ncu_lines.py $O/${TAG}_src_$w.csv 0.3 > $O/${TAG}_lines_$w.txt 2>&1
  rm -f $O/${TAG}_full_$w.ncu-rep
done
echo done
