#!/bin/bash
# Round evidence, run on a B200 through gpurun (see profiles/README_r01.md):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/collect_profiles.sh'
# Everything lands in gpurun_out/; tools/ncu_summary.py turns the reports into the text files under profiles/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_r01.txt
# 1. the bench line (our arm, then the reference arm on the host cores)
python bench.py > gpurun_out/bench_r01.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r01.log 2>&1
# 2. launch list of the bench command (one full step after the warm-up step; device time per launch).  Slow under ncu
#    (~13 min for the ~11 500 launches of a step): run it once per round.
ncu --metrics gpu__time_duration.sum --clock-control none -s 11500 -c 11500 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench_r01.log 2>&1
# 3. full-set captures of the kernels (one launch each; never a bench value)
ncu --set full --clock-control none --import-source on -k regex:k_poa_strip$ -s 50 -c 1 -o gpurun_out/prof_poa_strip_r01 \
    python tools/poa_bench.py --clusters 1200 --iters 1 --opt poa_units=2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_poa_strip_traceback -s 50 -c 1 -o gpurun_out/prof_poa_tb_r01 \
    python tools/poa_bench.py --clusters 1200 --iters 1 --opt poa_units=2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_poa_graph_fold -s 100 -c 1 -o gpurun_out/prof_fold_r01 \
    python tools/poa_bench.py --clusters 600 --iters 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_bv_scan -c 2 -o gpurun_out/prof_bv_r01 \
    python tools/bv_stream_bench.py --genes 8000 --seeds 1,512 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_join_count|k_pair_heavy|k_extract_smem" -s 40 -c 6 -o gpurun_out/prof_cluster_r01 \
    python bench.py --steps 1 --warmup 0 --no-correct --no-cpu-baseline > /dev/null 2>&1
# 4. the streaming / tiled regimes of the bitvector scan and the POA kernel alone (config-4 shape)
python tools/bv_stream_bench.py --genes 8000 > gpurun_out/bv_stream_r01.jsonl 2>&1
python tools/poa_bench.py --clusters 1200 --iters 2 > gpurun_out/poa_bench_r01.jsonl 2>&1
python tools/poa_bench.py --clusters 1200 --iters 1 --opt poa_kernel=1 > gpurun_out/poa_bench_int32_r01.jsonl 2>&1
ls -la gpurun_out
# 5. multi-GPU lines are taken separately (one process per GPU, never under ncu):
#   gpurun --gpus 4 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
#       --master-port 29511 bench.py --gpus 4 --steps 1 --warmup 1 --no-cpu-baseline'
