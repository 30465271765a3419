#!/bin/bash
# Round-2 evidence, part 3: parity suite on the final kernels, bench lines, bitvector-scan occupancy A/B, sanitizer logs.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_call3_r02.sh'
set -x
O=gpurun_out
mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > $O/pytest_gpu_r02c.log 2>&1
timeout 300 python bench.py --genes 2000 --steps 2 --warmup 1 --no-cpu-baseline > $O/b3_100k.json 2> $O/b3_100k.err
RATTLE_B200_THREADS=4 timeout 300 taskset -c 0-3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/b3_4cores.json 2> $O/b3_4cores.err
RTL_TRACE=1 timeout 300 python tools/poa_bench.py --clusters 1200 --iters 2 > $O/b3_poa_bench.jsonl 2> $O/b3_poa_bench.err
timeout 200 python tools/bv_stream_bench.py --genes 8000 --seeds 1,2,4,8,32,512 > $O/b3_bv_stream_minb4.jsonl 2>&1
timeout 200 python tools/bv_stream_bench.py --genes 8000 --seeds 1,2,4,8,32,512 --kernel 3 > $O/b3_bv_stream_minb3.jsonl 2>&1
# compute-sanitizer on the kernels that synchronise by hand (mailboxes between the DP warps, mbarrier seed tile) and on the vote kernels
S="tests/test_poa_gpu.py::test_correct_reads_matches_reference tests/test_poa_gpu.py::test_poa_golden_msa"
C="tests/test_cluster_gpu.py::test_cluster_reads_matches_oracle"
(timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $S -m gpu -q -x 2>&1 | tail -25) > $O/sanitizer_memcheck_poa_r02.log 2>&1
(timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $C -m gpu -q -x 2>&1 | tail -25) > $O/sanitizer_memcheck_cluster_r02.log 2>&1
(timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_poa_gpu.py::test_correct_reads_matches_reference -m gpu -q -x 2>&1 | tail -40) > $O/sanitizer_racecheck_poa_r02.log 2>&1
(timeout 300 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest "tests/test_cluster_gpu.py::test_cluster_reads_matches_oracle[512-False]" -m gpu -q -x 2>&1 | tail -40) > $O/sanitizer_racecheck_cluster_r02.log 2>&1
ls -la $O | tail -20
