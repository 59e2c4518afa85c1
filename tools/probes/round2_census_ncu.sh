for v in local global 5; do
  SCB_PREPARE_CENSUS=$v timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'spot_prepare|strip_fill|tile_scan' -c 9 --csv --log-file gpurun_out/r2r_launches_$v.csv python bench.py --resident-only --steps 2 --warmup 1 --frames-per-step 16 > /dev/null 2>&1
done
