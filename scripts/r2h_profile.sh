set -x
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2h_launches_vit_l16_b384.csv python scripts/profile_region.py vit_l16 > gpurun_out/r2h_prof_a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tn_kernel --launch-skip 1 --launch-count 4 -o gpurun_out/r2h_gemm4_b384 -f python scripts/profile_region.py vit_l16 > gpurun_out/r2h_prof_b.log 2>&1
python -m pytest tests/test_vit_gpu.py -q -m gpu -s -k "384 or 192" 2>&1 | grep -i "relative error\|passed\|failed"
