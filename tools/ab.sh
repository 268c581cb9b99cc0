#!/bin/bash
# A/B harness for gpurun: each line "NAME ENV..." runs a short resident-frames bench and prints one summary row.
run() {
  name=$1; shift
  out=$(env "$@" timeout 300 python bench.py --batch ${BATCH:-512} --steps ${STEPS:-6} --warmup 2 --no-cpu-baseline ${EXTRA:---no-breakdown} 2>&1 | tail -1)
  echo "$out" | python -c "
import sys, json
name = sys.argv[1]
try:
    d = json.loads(sys.stdin.read())
    k = d['kernel_ms_per_step']
    s = '%-28s value %.3e e2e %.3e  k2 %.2f ms  k3 %.2f ms  step %.2f ms' % (name, d['value'], d['e2e']['value'], k['k2_scan'], k['k3_cascade'], d['ms_per_step'])
    if 'by_distribution' in d:
        s += '  ' + ' '.join('%s %.2e' % (a, b['windows_per_s']) for a, b in d['by_distribution'].items())
    print(s)
except Exception as e:
    print(name, 'FAILED', e)
" "$name"
}
