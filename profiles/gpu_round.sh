#!/bin/bash
# One GPU visit: parity tests, bench line, phase times, ncu launch list, ncu full captures.
#   gpurun --timeout 900 -- 'bash profiles/gpu_round.sh TAG [quick]'
# Everything lands in gpurun_out/TAG_*; summaries worth keeping are copied to profiles/rNN/ by hand.
TAG=${1:-run}
QUICK=${2:-}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1
tail -3 $O/${TAG}_pytest_gpu.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python - <<EOF
import json
try:
    d = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("bench:", d["ms_per_step"], "ms/step", d["value"], d["unit"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"],
          "cpu", d["cpu_baseline"]["value"], "extra", d.get("extra"))
except Exception as e:
    print("bench failed", e)
    print(open("$O/${TAG}_bench.err").read()[-2000:])
EOF
python profiles/phase_times.py cfg2 cfg2_sparse > $O/${TAG}_phase_times.log 2>&1
cat $O/${TAG}_phase_times.log
[ -n "$QUICK" ] && exit 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-extra > $O/${TAG}_bench_under_ncu.log 2>&1
for w in cfg2 cfg2_sparse; do
  ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 4 -c 1 -f -o $O/${TAG}_full_$w \
      python profiles/run_profile.py --workload $w > $O/${TAG}_ncu_$w.log 2>&1
  ncu -i $O/${TAG}_full_$w.ncu-rep --page raw --csv > $O/${TAG}_raw_$w.csv 2>/dev/null
  python profiles/ncu_summary.py < $O/${TAG}_raw_$w.csv > $O/${TAG}_summary_$w.txt 2>&1
  head -40 $O/${TAG}_summary_$w.txt
done
