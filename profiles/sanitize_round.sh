set -x
cd $GRAFT_REPO_ROOT
PHASE_N=32 python profiles/phase_times.py cfg2 cfg2_sparse 2>&1 | tail -24
PHASE_N=148 python profiles/phase_times.py cfg2 2>&1 | tail -12
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "fused_vs_golden or nms_vs_oracle or target_loss_vs_golden or large_logits or decode_head_vs_golden or one_class or coco80 or cfg1" > gpurun_out/sanitizer_memcheck.log 2>&1; echo memcheck rc=$?; tail -5 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -x -q -k "fused_vs_golden or target_loss_vs_golden or nms_ties" > gpurun_out/sanitizer_racecheck.log 2>&1; echo racecheck rc=$?; tail -8 gpurun_out/sanitizer_racecheck.log
