M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
for w in mil_deploy mil_train virchow2 jpeg; do
  timeout 600 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2i_launches_$w.csv python scripts/profile_region.py $w > gpurun_out/r2i_prof_$w.log 2>&1
done
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
