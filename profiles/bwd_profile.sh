O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:target_loss_backward -s 13 -c 1 -f -o $O/bwd4 python profiles/loss_time.py 512 100 > $O/bwd4.log 2>&1
ncu -i $O/bwd4.ncu-rep --page raw --csv > $O/bwd4_raw.csv 2>/dev/null
python profiles/ncu_summary.py < $O/bwd4_raw.csv > $O/bwd4_summary.txt 2>&1
grep -E "gpu__time_duration|grid_size|issue_active.avg|stalled_barrier|long_scoreboard|inst_executed.sum|stalled_wait|short_scoreboard" $O/bwd4_summary.txt
ncu -i $O/bwd4.ncu-rep --page source --print-source cuda,sass --csv > $O/bwd4_src.csv 2>/dev/null
python profiles/ncu_lines.py $O/bwd4_src.csv 2.0 | head -40
