#!/bin/bash
# One GPU-box pass: tests, smoke, precision table, both bench arms.  Usage: tools/gpu_check.sh <tag> [steps...]
# Everything lands under gpurun_out/<tag>_*.  Steps: tests smoke precision bench ref ncu launches
TAG=${1:-run}; shift
STEPS=${@:-tests smoke bench}
mkdir -p gpurun_out
for s in $STEPS; do
  case $s in
    tests) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log;;
    smoke) timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/${TAG}_smoke.log;;
    precision) timeout 600 python tools/precision_stress.py gpurun_out/${TAG}_precision_stress.json > gpurun_out/${TAG}_precision_stress.md 2> gpurun_out/${TAG}_precision_stress.err; echo "precision rc=$?"; cat gpurun_out/${TAG}_precision_stress.md;;
    bench) timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench.err; cut -c1-700 gpurun_out/${TAG}_bench.json;;
    benchq) timeout 900 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench.err; cut -c1-700 gpurun_out/${TAG}_bench.json;;
    ref) timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_ref.json;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_infer4096.csv python tools/profile_forward.py 4096 infer bf16 > /dev/null 2>&1; echo "launches rc=$?";
              timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_train64.csv python tools/profile_forward.py 64 train ${TRAIN_PREC:-fp32} > /dev/null 2>&1; echo "launches-train rc=$?";;
    ncu) for ks in reg_layer_fused:1 sqa_fused:4 row_tail_fused:4; do   # kernel:launches to skip (2nd pass of profile_forward.py)
           k=${ks%%:*}; skip=${ks##*:}
           timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${TAG}_ncu_$k python tools/profile_forward.py 4096 infer bf16 > gpurun_out/${TAG}_ncu_$k.log 2>&1; echo "ncu $k rc=$?";
         done;;
  esac
done
