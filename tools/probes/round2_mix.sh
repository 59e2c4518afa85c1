set -x
mkdir -p gpurun_out
SCB_RENDER_COPY=mix timeout 300 python -m pytest tests/test_gpu_render.py tests/test_gpu_fullsize.py tests/test_gpu_geometries.py tests/test_gpu_movie.py -x -q 2>&1 | tail -5
for v in "SCB_RENDER_COPY=mix" "SCB_RENDER_COPY=tma"; do
  f=$(echo $v | tr ' =' '__')
  env $v timeout 300 python bench.py --resident-only --steps 4 > gpurun_out/r2h_bench_$f.json 2> gpurun_out/r2h_bench_$f.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2h_bench_$f.json").read().strip().splitlines()[-1])
print("VARIANT $v: frames/s %.0f render ms %.4f step ms %.3f checksum %.6f" % (d["value"], d["render_ms_per_launch"], d["ms_per_step"], d["frame_checksum_mean_adc"]))
P
done
