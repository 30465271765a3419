#!/bin/bash
# Round-2 evidence, part 1 (run on a B200 through gpurun):
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/gpu_call1_r02.sh'
# parity suite, bench lines, host-phase traces, launch list and the full-set ncu captures.  Everything lands in
# gpurun_out/; tools/ncu_summary.py turns the reports into the text files under profiles/.
# Every step has its own timeout: a hung step must not eat the GPU budget.
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi_r02.txt
nproc > $O/nproc_r02.txt
# 1. the GPU parity suite
(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > $O/pytest_gpu_r02.log 2>&1
# 2. the bench line (our arm, N=1: 125 k reads) and configs[1] itself (100 k reads) with the library's phase trace
timeout 600 python bench.py > $O/bench_n1_r02.json 2> $O/bench_n1_r02.err
RTL_TRACE=1 timeout 300 python bench.py --genes 2000 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_n1_100k_r02.json 2> $O/bench_n1_100k_r02.err
# 3. the host share of one rank of an 8-rank node (4 cores per rank) reproduced on one GPU
RTL_TRACE=1 RATTLE_B200_THREADS=4 timeout 300 taskset -c 0-3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $O/bench_n1_4cores_r02.json 2> $O/bench_n1_4cores_r02.err
# 4. launch list of the bench command (one full step after a warm-up step; device time per launch, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_r02.csv \
    python bench.py --genes 2000 --steps 1 --warmup 1 --no-cpu-baseline > $O/launches_bench_r02.log 2>&1
# 5. full-set captures (one launch each; never a bench value); small arena so that ncu's save/restore stays cheap
timeout 420 ncu --set full --clock-control none --import-source on -k regex:k_poa_chain -s 1 -c 1 -o $O/prof_poa_chain_r02 \
    python tools/poa_bench.py --clusters 300 --iters 1 --opt poa_units=1 --opt poa_arena_mb=12000 > $O/ncu_poa_chain.log 2>&1
timeout 240 ncu --set full --clock-control none -k regex:k_bv_scan -c 2 -o $O/prof_bv_scan_r02 \
    python tools/bv_stream_bench.py --genes 8000 --seeds 1,512 --reps 1 > /dev/null 2>&1
timeout 240 ncu --set full --clock-control none -k regex:k_bv_stream -c 1 -o $O/prof_bv_stream_r02 \
    python tools/bv_stream_bench.py --genes 8000 --seeds 1 --reps 1 --kernel 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_join_count|k_pair_heavy|k_extract_smem" -s 40 -c 6 -o $O/prof_cluster_r02 \
    python bench.py --genes 2000 --steps 1 --warmup 0 --no-correct --no-cpu-baseline > /dev/null 2>&1
# 6. the two regimes of the bitvector scan (default kernel, then the bulk-copy ring kernel), POA alone (config-4 shape)
timeout 200 python tools/bv_stream_bench.py --genes 8000 > $O/bv_stream_r02.jsonl 2>&1
timeout 200 python tools/bv_stream_bench.py --genes 8000 --seeds 1,2,4,8,16 --kernel 2 > $O/bv_stream_ring_r02.jsonl 2>&1
RTL_TRACE=1 timeout 300 python tools/poa_bench.py --clusters 1200 --iters 2 > $O/poa_bench_r02.jsonl 2> $O/poa_bench_r02.err
timeout 300 python tools/poa_bench.py --clusters 1200 --iters 1 --opt poa_device_chain=0 > $O/poa_bench_hostpath_r02.jsonl 2>&1
ls -la $O | tail -40
