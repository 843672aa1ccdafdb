#!/bin/bash
# What the driver runs at round end, plus the evidence of this session: GPU tests, smoke, the default bench line,
# memcheck of the table-build kernels, one ncu --set full capture (with source) of the headline kernel.
# usage (on the GPU box): bash tools/round_check.sh <tag>
tag=${1:-r01b}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
( time python bench.py ) > gpurun_out/${tag}_bench_headline.json 2> gpurun_out/${tag}_bench_headline.err
head -c 400 gpurun_out/${tag}_bench_headline.json; echo
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -x -q -m gpu \
  "tests/test_gpu_parity.py::test_bokeh_tables_built_on_the_gpu_equal_the_oracle_tables" 2>&1 \
  | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -5 > gpurun_out/${tag}_sanitizer_bokeh.txt
cat gpurun_out/${tag}_sanitizer_bokeh.txt
ncu --set full --clock-control none --import-source on -k regex:kolb_pool2 -s 2 -c 1 -o gpurun_out/${tag}_ncu_pool2 \
    python bench.py --spp 4 --steps 1 --warmup 2 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_pool2.log 2>&1
ls -la gpurun_out | head -30
