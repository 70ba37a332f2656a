#!/bin/bash
# compute-sanitizer memcheck over the tests of the kernels added late in round 1 (channels-last decode, large-image
# path, compact all-gather rows), racecheck over the large-image path (spin flags + staged pair blocks).
#   gpurun --timeout 1500 -- 'bash profiles/sanitize_round2.sh'
O=gpurun_out
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q \
  -k "channels_last or large_image or host_pipeline or compact or unsupported" \
  > $O/sanitizer_memcheck2.log 2>&1; echo memcheck rc=$?; tail -4 $O/sanitizer_memcheck2.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -x -q \
  -k "large_image_path_equals_fused" \
  > $O/sanitizer_racecheck2.log 2>&1; echo racecheck rc=$?; grep -E "Race reported|RACECHECK SUMMARY|hazard" $O/sanitizer_racecheck2.log | sort | uniq -c | sort -rn | head -12
