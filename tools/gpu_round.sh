#!/bin/bash
# One gpurun call's worth of verification for the current tree: GPU parity tests, smoke, the bench line,
# the reference arm, the ncu launch list, one `--set full` capture of the three kernels (digested on the
# box), and a short A/B of k2_scan's phase schedule.  Everything lands in gpurun_out/<tag>_*.
#   usage (on the GPU box, from the repo root): [NCU=0] [AB=1] bash tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-round}
O=gpurun_out   # (tools/ab.sh uses $out itself)
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${tag}_gpu.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -x -q -m gpu > $O/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> $O/${tag}_pytest.log
  tail -5 $O/${tag}_pytest.log
fi
timeout 300 python __graft_entry__.py smoke > $O/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> $O/${tag}_smoke.log
tail -2 $O/${tag}_smoke.log
timeout 600 python bench.py > $O/${tag}_bench.json 2> $O/${tag}_bench.err
tail -c 1500 $O/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${tag}_bench_ref.json 2>> $O/${tag}_bench.err
cat $O/${tag}_bench_ref.json
if [ "${NCU:-1}" = "1" ]; then
# launch list of the same command (per-launch times are cold-cache and serialised: only the shares count)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-breakdown > $O/${tag}_ncu_bench.log 2>&1
grep -c k2_scan $O/${tag}_launches.csv
# full capture: the kernels of ONE 512-frame step (the stage kernels run from 6e7 candidate windows up) (k2_scan, k3_regress, 4 x (k3_walk, k3_regress), k3_emit), resident frames;
# the traced instantiations the bench uses for its cart statistics are left out by name
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'k2_scan<\(int\)4, \(bool\)0, \(bool\)0>|k3_walk<\(bool\)0|k3_regress|k3_emit<\(bool\)0' -s 11 -c 11 -f -o /tmp/${tag}_full \
  python bench.py --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-breakdown > $O/${tag}_ncu_full.log 2>&1
python tools/ncu_digest.py /tmp/${tag}_full.ncu-rep $O/${tag}_full 512 >> $O/${tag}_ncu_full.log 2>&1
cp $O/${tag}_full_k2_capture.json $O/k2_capture.json 2>/dev/null   # -> profiles/k2_capture.json (bench.py reads it)
ls -la /tmp/${tag}_full.ncu-rep >> $O/${tag}_ncu_full.log 2>&1
fi
if [ "${AB:-0}" = "1" ]; then
# A/B of the cascade kernels (resident frames, 6 steps each)
source tools/ab.sh
{
run default
run k3_cascade_for_stages_ge_1 JDA_B200_NO_STAGE_KERNELS=1
run k3_stage0_regression JDA_B200_OLD_REGRESS=1
run round1_k3 JDA_B200_NO_STAGE_KERNELS=1 JDA_B200_OLD_REGRESS=1
} > $O/${tag}_ab.txt 2>&1
cat $O/${tag}_ab.txt
fi
