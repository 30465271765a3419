#!/bin/bash
# Round-2 experiments: narrow CTAs for the device-resident chains (RATTLE_B200_NARROW=1).
set -x
O=gpurun_out
mkdir -p $O
(time RATTLE_B200_NARROW=1 timeout 600 python -m pytest tests/test_poa_gpu.py tests/test_golden_big_gpu.py -m gpu -q 2>&1 | tail -8) > $O/pytest_gpu_narrow.log 2>&1
timeout 300 python bench.py --genes 2000 --steps 2 --warmup 1 --no-cpu-baseline > $O/b4_100k.json 2> $O/b4_100k.err
RATTLE_B200_NARROW=1 RTL_TRACE=1 timeout 300 python bench.py --genes 2000 --steps 2 --warmup 1 --no-cpu-baseline > $O/b4_100k_narrow.json 2> $O/b4_100k_narrow.err
RTL_TRACE=1 timeout 300 python tools/poa_bench.py --clusters 1200 --iters 2 > $O/b4_poa_bench.jsonl 2> $O/b4_poa_bench.err
RATTLE_B200_NARROW=1 RTL_TRACE=1 timeout 300 python tools/poa_bench.py --clusters 1200 --iters 2 > $O/b4_poa_bench_narrow.jsonl 2> $O/b4_poa_bench_narrow.err
ls -la $O | tail -8
