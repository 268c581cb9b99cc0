#!/bin/bash
# SASS evidence for the TMA / mbarrier / cp.async claims: per kernel, the count of the mnemonics that prove them
# (B200_PROFILING.md "What proves a Blackwell-native kernel") plus the lines themselves.  Runs without a GPU.
#   usage: bash tools/sass_evidence.sh > profiles/r2_k2_scan.sass.txt
so=${1:-jda_b200/libjda_b200.so}
echo "# cuobjdump -sass $so  (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), sm_100a)"
cuobjdump -sass "$so" | awk '
  /Function :/ { fn=$3; next }
  /UTMALDG|UTMASTG|UBLKCP|SYNCS|LDGSTS|UTC.MMA|LDTM|STTM/ {
    m=$0; sub(/^[ \t]*\/\*[0-9a-f]+\*\/[ \t]*/, "", m); sub(/[ \t]*\/\*.*$/, "", m);
    n=m; sub(/[ .].*$/, "", n); if (n ~ /^@/) { split(m, a, " "); n=a[2]; sub(/\..*$/, "", n) }
    cnt[fn "\t" n]++; if (seen[fn "\t" m]++ == 0) ex[fn]=ex[fn] "\n      " m
  }
  END { for (k in cnt) print k "\t" cnt[k] | "sort"; close("sort"); print ""; for (f in ex) print f ":" ex[f] }'
