# Evidence of the ragged-plan build: launch lists (dense with the CHROMO_F_DENSE hint, ragged with the plan), ncu --set full of
# the Regulation kernel (token-class tiles) and of the Pairwise single-query attention (key windows) on the ragged sweep.
TAG=${1:-r03}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_bf16_infer18955.csv python tools/profile_forward.py 18955 infer bf16 > /dev/null 2>&1; echo "launches dense rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_bf16_infer18955_ragged.csv python tools/profile_forward.py 18955 infer bf16 ragged > /dev/null 2>&1; echo "launches ragged rc=$?"
for ks in reg_layer_fused:1 sqa_fused:4 row_tail_fused:4; do
  k=${ks%%:*}; skip=${ks##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${TAG}_ncu_ragged_$k python tools/profile_forward.py 18955 infer bf16 ragged > gpurun_out/${TAG}_ncu_ragged_$k.log 2>&1; echo "ncu $k rc=$?"
done
