#!/bin/bash
# shared-memory wavefronts / bank-conflict replays / duration of one k2_scan launch (256 resident frames) under an env variant
#   usage: [DIST=facemix] tools/ncu_wavefronts.sh NAME [ENV=VALUE ...]
name=$1; shift
log=/tmp/ncu_wf_$name.csv
env "$@" timeout 600 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,gpu__time_duration.sum,smsp__inst_executed.sum \
  --clock-control none -k regex:k2_scan -s 3 -c 1 --csv --log-file $log python bench.py --batch 256 --dist ${DIST:-facemix} --steps 2 --warmup 1 --no-cpu-baseline --no-breakdown > /dev/null 2>&1
python - "$name" "$log" <<'PY'
import csv, sys
name, log = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(log)) if len(r) > 12 and "k2_scan" in r[4]]
print(name, " ".join("%s=%s" % (r[12].split("__")[-1][:34], r[14]) for r in rows))
PY
