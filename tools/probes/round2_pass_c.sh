# Round-2 pass C (one B200): block route of generate_images -- API / movie / render tests, then the weak bench
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_api.py tests/test_gpu_movie.py tests/test_gpu_render.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -15 > gpurun_out/r2c_tests.log; tail -15 gpurun_out/r2c_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_weak.json 2> gpurun_out/r2c_bench_weak.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c_bench_weak.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", json.dumps(d["e2e"])[:900])
P
tail -c 800 gpurun_out/r2c_bench_weak.err
