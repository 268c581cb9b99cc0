#!/bin/bash
# per-level ncu counters of k2_scan: one level per launch (tools/level_probe.py --once), light metric set.
#   usage (GPU box): bash tools/level_ncu.sh <tag> [ENV=VALUE ...]
tag=$1; shift
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
env "$@" timeout 900 ncu --metrics $M --clock-control none -k regex:"k2_scan|k1_planes" --csv --log-file /tmp/lvncu_$tag.csv \
  python tools/level_probe.py --once > gpurun_out/${tag}_level_probe_under_ncu.txt 2>&1
python - /tmp/lvncu_$tag.csv > gpurun_out/${tag}_level_ncu.txt <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
by, name = {}, {}
for r in rows:
    by.setdefault(int(r[0]), {})[r[12]] = r[14]
    name[int(r[0])] = r[4][5:14]
short = lambda k: k.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '').replace('l1tex__', '').replace('.sum', '').replace('.avg.pct_of_peak_sustained_active', '%').replace('.pct_of_peak_sustained_elapsed', '%')
for i, (k, d) in enumerate(sorted(by.items())):
    print('launch %2d %s ' % (i, name[k]) + '  '.join('%s=%s' % (short(a), b) for a, b in d.items()))
PY
cat gpurun_out/${tag}_level_ncu.txt
