#!/bin/bash
# Multi-GPU measurements on ONE box (run under `gpurun --gpus 8`): the 2-GPU NCCL parity test, then config 4 and
# config 5 at 1 / 2 / 4 / 8 ranks and the headline workload at 8.  One JSON line per run in gpurun_out/<tag>_<workload>_n<N>.json.
tag=${1:-multi}
maxn=${2:-8}   # GPUs on the box
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/${tag}_gpus.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $O/${tag}_pytest_multi.log 2>&1; tail -2 $O/${tag}_pytest_multi.log
port=29600
run() {  # workload N extra-args...
  w=$1; n=$2; shift 2
  port=$((port + 1))
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --workload $w "$@" > $O/${tag}_${w}_n$n.json 2> $O/${tag}_${w}_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --workload $w "$@" > $O/${tag}_${w}_n$n.json 2> $O/${tag}_${w}_n$n.err
  fi
  python - $O/${tag}_${w}_n$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    chk = d.get("gather_check") or d.get("gather_check_last_batch") or {}
    print("%s n=%d value %.3e e2e %.3e ms/step %.1f check %s" % (d["config"]["workload"], d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], chk))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
# (GPU-minutes are charged per GPU of the box: CFG4_NS / CFG5_NS / VGA_NS trim the sweep)
for n in ${CFG4_NS-1 2 4 8}; do [ $n -le $maxn ] && run cfg4 $n --steps 3 --warmup 3; done
for n in ${CFG5_NS-8 4 2 1}; do [ $n -le $maxn ] && run cfg5 $n --steps 1 --warmup 3; done
for n in ${VGA_NS-$maxn 2}; do [ $n -le $maxn ] && run vga $n --steps 10 --warmup 3 --no-breakdown; done
true
