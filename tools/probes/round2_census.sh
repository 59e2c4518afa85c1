# shared-memory census of spot_prepare against the global-atomic one: parity tests, resident bench, kernel times
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_geometries.py tests/test_gpu_fullsize.py tests/test_gpu_movie.py tests/test_gpu_api.py tests/test_gpu_gaussian_tc.py -x -q 2>&1 | tail -3
for v in local global 5; do
  SCB_PREPARE_CENSUS=$v timeout 300 python bench.py --resident-only --steps 6 > gpurun_out/r2r_bench_$v.json 2> gpurun_out/r2r_bench_$v.err
  SCB_PREPARE_CENSUS=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spot_prepare|strip_fill|tile_scan' -s 6 -c 6 --csv --log-file gpurun_out/r2r_launches_$v.csv python bench.py --resident-only --steps 1 --warmup 1 --frames-per-step 32 > /dev/null 2>&1
done
python - <<'P'
import json, csv
for v in ("local", "global", "5"):
    d = json.loads(open("gpurun_out/r2r_bench_%s.json" % v).read().strip().splitlines()[-1])
    print("CENSUS %s: frames/s %.0f render ms/launch %.4f step ms %.3f checksum %.6f" % (v, d["value"], d["render_ms_per_launch"], d["ms_per_step"], d["frame_checksum_mean_adc"]))
    rows = [r for r in csv.reader(open("gpurun_out/r2r_launches_%s.csv" % v)) if len(r) > 5 and r[0].isdigit()]
    for r in rows:
        print("   ", r[4][:60], r[-1], r[-2])
P
