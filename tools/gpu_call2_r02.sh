#!/bin/bash
# Round-2 evidence, part 2: device-side vote, traceback prefetch, CTA-per-SM experiments.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_call2_r02.sh'
set -x
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > $O/pytest_gpu_r02b.log 2>&1
RTL_TRACE=1 timeout 300 python bench.py --genes 2000 --steps 2 --warmup 1 --no-cpu-baseline > $O/b2_100k_vote.json 2> $O/b2_100k_vote.err
timeout 300 python bench.py --genes 2000 --steps 2 --warmup 1 --no-cpu-baseline --opt poa_device_vote=0 > $O/b2_100k_hostvote.json 2> $O/b2_100k_hostvote.err
RATTLE_B200_CTAS5=1 timeout 300 python bench.py --genes 2000 --steps 2 --warmup 1 --no-cpu-baseline > $O/b2_100k_vote_ctas5.json 2> $O/b2_100k_vote_ctas5.err
RTL_TRACE=1 RATTLE_B200_THREADS=4 timeout 300 taskset -c 0-3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/b2_4cores_vote.json 2> $O/b2_4cores_vote.err
RTL_TRACE=1 timeout 300 python tools/poa_bench.py --clusters 1200 --iters 2 > $O/b2_poa_bench.jsonl 2> $O/b2_poa_bench.err
RATTLE_B200_CTAS4W8=1 timeout 300 python tools/poa_bench.py --clusters 1200 --iters 2 > $O/b2_poa_bench_ctas4w8.jsonl 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k regex:k_poa_chain -c 1 -o $O/prof_poa_chain_r02b \
    python tools/poa_bench.py --clusters 600 --iters 1 --opt poa_units=1 --opt poa_arena_mb=24000 > $O/ncu_poa_chain_b.log 2>&1
ls -la $O | tail -20
