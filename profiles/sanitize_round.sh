#!/bin/bash
# compute-sanitizer over the GPU parity tests (memcheck: all kernels; racecheck: the shared-memory heavy ones).
#   gpurun --timeout 1500 -- 'bash profiles/sanitize_round.sh'
O=gpurun_out
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q \
  -k "golden or nms_vs_oracle or large_logits or one_class or coco80 or cfg1 or backward or map or seg or box_ciou or back_to_back or chain" \
  > $O/sanitizer_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 $O/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -x -q \
  -k "fused_vs_golden or target_loss_vs_golden or nms_ties or backward_vs_reference or map_vs_reference" \
  > $O/sanitizer_racecheck.log 2>&1; echo racecheck rc=$?; grep -E "Race reported|RACECHECK SUMMARY" $O/sanitizer_racecheck.log | sort | uniq -c | sort -rn | head -12
