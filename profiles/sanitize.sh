#!/bin/bash
# compute-sanitizer over small-shape parity tests: memcheck, racecheck, synccheck.
# Usage (on a GPU box, from the repo root): bash profiles/sanitize.sh OUTDIR
# The selections keep to small shapes: the tools slow a kernel down 10-100x.
out=${1:-gpurun_out}
sel_setops='kat or seam_random or structured_overlaps or unaligned or wide_key or calling_convention'
sel_ops='fuzz or slab_ignores or ctor or transpose_shape or long_segments'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 77 \
      python -m pytest tests/test_gpu_setops.py -x -q -k "$sel_setops" > $out/sanitize_${tool}_setops.log 2>&1
  echo "setops rc=$?" | tee -a $out/sanitize_${tool}_setops.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 77 \
      python -m pytest tests/test_gpu_ops.py tests/test_gpu_sort.py -x -q -k "$sel_ops" > $out/sanitize_${tool}_ops.log 2>&1
  echo "ops+sort rc=$?" | tee -a $out/sanitize_${tool}_ops.log
done
