#!/bin/bash
# Round 2: compute-sanitizer over the GPU parity tests of the rebuilt fused kernel (fp16 prefilter, strip tasks with
# acquire / release hand-overs in shared memory), the new entry points and the loss kernel's prefilter.
#   gpurun --timeout 1800 -- 'bash profiles/sanitize_round3.sh'
O=gpurun_out
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "golden or nms_vs_oracle or large_logits or one_class or coco80 or cfg1 or tile_boundaries or batches or lazy or host_entry or channels_last or out_of_range or backward or map or seg or box_ciou or back_to_back or chain" \
  > $O/sanitizer_memcheck_r02.log 2>&1; echo memcheck rc=$?; tail -4 $O/sanitizer_memcheck_r02.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "fused_vs_golden or tile_boundaries or nms_ties or target_loss_vs_golden or cfg3_bdd or cfg5_416" \
  > $O/sanitizer_racecheck_r02.log 2>&1; echo racecheck rc=$?
grep -E "Race reported|RACECHECK SUMMARY|ERROR SUMMARY" $O/sanitizer_racecheck_r02.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -12
tail -3 $O/sanitizer_racecheck_r02.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer_gather_single_rank" \
  > $O/sanitizer_memcheck_gather_r02.log 2>&1; echo gather memcheck rc=$?; tail -3 $O/sanitizer_memcheck_gather_r02.log
