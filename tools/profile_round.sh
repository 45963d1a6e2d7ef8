TAG=r02t
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_infer4096.csv python tools/profile_forward.py 4096 infer bf16 > /dev/null 2>&1; echo "launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_train64_bf16.csv python tools/profile_forward.py 64 train bf16 > /dev/null 2>&1; echo "launches-train rc=$?"
for ks in reg_layer_fused:1 sqa_fused:4 row_tail_fused:4; do
  k=${ks%%:*}; skip=${ks##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${TAG}_ncu_$k python tools/profile_forward.py 4096 infer bf16 > gpurun_out/${TAG}_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_grouped -s 2 -c 1 -f -o gpurun_out/${TAG}_ncu_wgrad_grouped python tools/profile_forward.py 64 train bf16 > gpurun_out/${TAG}_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 230 -c 1 -f -o gpurun_out/${TAG}_ncu_tc_gemm python tools/profile_forward.py 64 train bf16 > gpurun_out/${TAG}_ncu_tc_gemm.log 2>&1; echo "ncu tc_gemm rc=$?"
