#!/bin/bash
# compute-sanitizer memcheck + racecheck over the tiled kernels (small frames: the GPU parity tests themselves).
# usage (on a GPU box): bash tools/sanitize.sh gpurun_out/r2_sanitize
out=${1:-gpurun_out/sanitize}
sel='stack or blur or resize or rotate or transitions'
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "$sel" -p no:cacheprovider > ${out}_${tool}_ops.log 2>&1
  echo "$tool ops rc=$?" | tee -a ${out}_summary.txt
  compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest tests/test_y4m.py -x -q -m gpu -k "kernel_matches_libswscale_small or quantises" -p no:cacheprovider > ${out}_${tool}_y4m.log 2>&1
  echo "$tool y4m rc=$?" | tee -a ${out}_summary.txt
done
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" ${out}_*_ops.log ${out}_*_y4m.log | tee -a ${out}_summary.txt
