set -x
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
for w in vit_l16 gigapath resize; do
  timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2g_launches_$w.csv python scripts/profile_region.py $w > gpurun_out/r2g_prof_$w.log 2>&1
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tn_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/r2g_fc1 -f python scripts/profile_region.py vit_l16 > gpurun_out/r2g_prof_fc1.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:resize_u8_kernel --launch-count 1 -o gpurun_out/r2g_resize -f python scripts/profile_region.py resize > gpurun_out/r2g_prof_resize.log 2>&1
ls -la gpurun_out/ | tail -8
