# quick check of a render change: parity tests of the render paths, then the resident bench
set -x
timeout 400 python -m pytest tests/test_gpu_render.py tests/test_gpu_geometries.py tests/test_gpu_fullsize.py tests/test_gpu_movie.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --resident-only --steps 6 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2q_bench.json").read().strip().splitlines()[-1])
print("QUICK: frames/s %.0f render ms/launch %.4f step ms %.3f checksum %.6f" % (d["value"], d["render_ms_per_launch"], d["ms_per_step"], d["frame_checksum_mean_adc"]))
P
