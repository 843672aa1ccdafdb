#!/bin/bash
# round 2, GPU call 4 (TWO GPUs): the NVLink gather (three transports) and the strong-split bench line at N = 2.
tag=r02d
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
( time timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jobs.py -x -q -m gpu -k "synth or streamed_job or census or differentials or rearm" ) > gpurun_out/${tag}_pytest_quick.log 2>&1
tail -4 gpurun_out/${tag}_pytest_quick.log
( time timeout 600 python -m pytest tests/test_gpu_jobs.py -x -q -m gpu -k "gather" -s ) > gpurun_out/${tag}_pytest_gather.log 2>&1
tail -6 gpurun_out/${tag}_pytest_gather.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/${tag}_bench_headline_2gpu.json 2> gpurun_out/${tag}_bench_headline_2gpu.err
head -c 1500 gpurun_out/${tag}_bench_headline_2gpu.json; echo
tail -5 gpurun_out/${tag}_bench_headline_2gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --workload config4 --no-e2e --transports fused ) > gpurun_out/${tag}_bench_config4_2gpu.json 2> gpurun_out/${tag}_bench_config4_2gpu.err
head -c 600 gpurun_out/${tag}_bench_config4_2gpu.json; echo
tail -3 gpurun_out/${tag}_bench_config4_2gpu.err
timeout 300 python bench.py --workload headline --stream --steps 5 --warmup 3 --no-cpu --no-e2e --census-rays 0 > gpurun_out/${tag}_bench_headline_streamed.json 2>> gpurun_out/${tag}.err
head -c 300 gpurun_out/${tag}_bench_headline_streamed.json; echo
