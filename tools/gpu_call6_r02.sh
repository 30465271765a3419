#!/bin/bash
# Round-2 evidence, final single-GPU pass: parity suite, bench lines (our arm and the reference arm), launch list,
# full-set ncu captures of the shipped kernels, racecheck details.
set -x
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > $O/pytest_gpu_final_r02.log 2>&1
timeout 600 python bench.py > $O/bench_n1_final_r02.json 2> $O/bench_n1_final_r02.err
timeout 300 python bench.py --genes 2000 --no-cpu-baseline > $O/bench_n1_100k_final_r02.json 2> $O/bench_n1_100k_final_r02.err
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > $O/bench_ref_final_r02.json 2> $O/bench_ref_final_r02.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_final_r02.csv \
    python bench.py --genes 2000 --steps 1 --warmup 1 --no-cpu-baseline > $O/launches_bench_final_r02.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k regex:k_poa_chain -c 1 -o $O/prof_poa_chain_final_r02 \
    python tools/poa_bench.py --clusters 600 --iters 1 --opt poa_units=1 --opt poa_arena_mb=24000 > $O/ncu_poa_chain_final.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_vote_rows|k_vote_cols|k_vote_apply|k_pack_bases|k_chain_msa_rows" -c 6 -o $O/prof_vote_final_r02 \
    python bench.py --genes 400 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
timeout 240 ncu --set full --clock-control none -k regex:k_bv_scan -c 2 -o $O/prof_bv_scan_final_r02 \
    python tools/bv_stream_bench.py --genes 8000 --seeds 1,512 --reps 1 > /dev/null 2>&1
(timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 12 python -m pytest tests/test_poa_gpu.py::test_correct_reads_matches_reference -m gpu -q -x 2>&1 | grep -v "Host Frame\|^=========$" | head -120) > $O/sanitizer_racecheck_poa_detail_r02.log 2>&1
ls -la $O | tail -12
