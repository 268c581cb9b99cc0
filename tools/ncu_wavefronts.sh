#!/bin/bash
# shared-memory wavefronts / bank-conflict replays / duration of one k2_scan launch (256 resident frames) under an env variant
#   usage: tools/ncu_wavefronts.sh NAME [ENV=VALUE ...]
name=$1; shift
env "$@" timeout 600 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,gpu__time_duration.sum,smsp__inst_executed.sum \
  --clock-control none -k regex:k2_scan -s 3 -c 1 --csv python bench.py --batch 256 --dist ${DIST:-facemix} --steps 2 --warmup 1 --no-cpu-baseline --no-breakdown 2>/dev/null \
  | grep -E "k2_scan" | awk -F'","' -v n="$name" '{gsub(/"/,"",$NF); printf "%s %s %s\n", n, $(NF-2), $NF}'
