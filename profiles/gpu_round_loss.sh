#!/bin/bash
# ncu --set full captures of the secondary kernels (loss forward / backward, seg head, mAP) -> gpurun_out/TAG_*
#   gpurun --timeout 900 -- 'bash profiles/gpu_round_loss.sh TAG'
TAG=${1:-loss}
O=gpurun_out
mkdir -p $O
cap() {  # name kernel-regex skip script args...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/${TAG}_full_$name "$@" > $O/${TAG}_ncu_$name.log 2>&1
  ncu -i $O/${TAG}_full_$name.ncu-rep --page raw --csv > $O/${TAG}_raw_$name.csv 2>/dev/null
  python profiles/ncu_summary.py < $O/${TAG}_raw_$name.csv > $O/${TAG}_summary_$name.txt 2>&1
  grep -E "Kernel|==|gpu__time_duration|dram__bytes|issue_active.avg.pct|registers_per_thread \[" $O/${TAG}_summary_$name.txt | head -8
}
cap target_loss_fwd_head1 'target_loss_kernel' 13 python profiles/loss_time.py 512 100
cap target_loss_bwd_head1 'target_loss_backward' 13 python profiles/loss_time.py 512 100
cap seg_loss 'seg_loss_kernel' 20 python profiles/seg_time.py
cap map_match 'map_match' 3 python profiles/map_time.py 4952 40
cap map_class 'map_class' 3 python profiles/map_time.py 4952 40
