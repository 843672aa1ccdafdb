#!/bin/bash
# memcheck + racecheck over tests that drive the persistent pool kernels across chunk boundaries (sample prefetch, slot
# pools, stacks) and the image-sampling tables.  usage (on the GPU box): bash tools/sanitize_pool.sh > gpurun_out/sanitizer_pool.txt
SEL="tests/test_gpu_parity.py::test_large_batch_spans_many_chunks_and_keeps_counters tests/test_gpu_parity.py::test_batch_boundaries_do_not_matter tests/test_gpu_parity.py::test_thin_lens_hex_bokeh tests/test_gpu_parity.py::test_guarded_kolb_no_lut_and_bokeh tests/test_gpu_parity.py::test_guarded_kolb_bokeh_image_sizes"
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Race|hazard" | head -8
done
