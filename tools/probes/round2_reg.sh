# Register-accumulator render: parity tests, then the measured variants against the shared-memory kernel (one B200)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render.py tests/test_gpu_geometries.py tests/test_gpu_fullsize.py tests/test_gpu_movie.py tests/test_gpu_api.py -x -q 2>&1 | tail -15 > gpurun_out/r2c_tests.log
tail -5 gpurun_out/r2c_tests.log
for v in "SCB_RENDER_OCC=0" "SCB_RENDER_OCC=1 SCB_RENDER_DIST=1" "SCB_RENDER_OCC=1 SCB_RENDER_DIST=2" "SCB_RENDER_OCC=2 SCB_RENDER_DIST=1" \
         "SCB_RENDER_OCC=3 SCB_RENDER_DIST=2" "SCB_RENDER_OCC=3 SCB_RENDER_DIST=3" "SCB_RENDER_OCC=3 SCB_RENDER_DIST=5" \
         "SCB_RENDER_OCC=4 SCB_RENDER_DIST=4" "SCB_RENDER_OCC=4 SCB_RENDER_DIST=6" "SCB_RENDER_REG=tensor" "SCB_RENDER_PATH=smem"; do
  f=$(echo $v | tr ' =' '__')
  env $v python bench.py --resident-only --steps 4 > gpurun_out/r2c_bench_$f.json 2> gpurun_out/r2c_bench_$f.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2c_bench_$f.json").read().strip().splitlines()[-1])
print("VARIANT $v: frames/s %.0f render ms %.4f" % (d["value"], d["render_ms_per_launch"]))
P
done
