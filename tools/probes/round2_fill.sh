# strip_fill variants: list positions requested up front (3 / 4 / 2 CTAs per SM) against the strip-by-strip loads
set -x
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_geometries.py tests/test_gpu_fullsize.py tests/test_gpu_movie.py -x -q 2>&1 | tail -3
for v in 3 old 4 2; do
  SCB_FILL_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes.sum --clock-control none -k regex:'spot_prepare|strip_fill' -c 6 --csv --log-file gpurun_out/r2t_launches_$v.csv python bench.py --resident-only --steps 2 --warmup 1 --frames-per-step 16 > /dev/null 2>&1
done
timeout 300 python bench.py --resident-only --steps 6 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
tail -c 600 gpurun_out/r2t_bench.json
