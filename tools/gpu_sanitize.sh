#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on shared memory) over a small slice of the GPU tests.
O=gpurun_out; mkdir -p $O; tag=${1:-san}
K1='known_answer or mixed_sizes_edge or edge_cases or tile_origins'
K2='known_answer or refusals or training_snapshot'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "$K1" > $O/${tag}_memcheck_c.log 2>&1; echo "exit $?" >> $O/${tag}_memcheck_c.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_cpp_path.py -x -q -k "$K2" > $O/${tag}_memcheck_cpp.log 2>&1; echo "exit $?" >> $O/${tag}_memcheck_cpp.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "known_answer" > $O/${tag}_racecheck.log 2>&1; echo "exit $?" >> $O/${tag}_racecheck.log
# round 2: the stage kernels (k3_walk / k3_regress / k3_emit, truncated cascades included) and the jdaDetect combiner
K3='(stage_kernels and t2_k101) or (cart_granular and 6-t4_k270) or coalesced or (trace_throughput and faces and tma)'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "$K3" > $O/${tag}_memcheck_stages.log 2>&1; echo "exit $?" >> $O/${tag}_memcheck_stages.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "(stage_kernels and t2_k101) or (trace_throughput and faces and tma)" > $O/${tag}_racecheck_stages.log 2>&1; echo "exit $?" >> $O/${tag}_racecheck_stages.log
for f in memcheck_c memcheck_cpp racecheck memcheck_stages racecheck_stages; do echo "== $f"; grep -E "ERROR SUMMARY|passed|failed|exit|Invalid|hazard" $O/${tag}_$f.log | head -12; done
