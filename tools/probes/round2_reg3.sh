set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_render.py -x -q -k "register_kernels" 2>&1 | tail -3
for v in "SCB_RENDER_PATH=ldg" "SCB_RENDER_PATH=smem"; do
  f=$(echo $v | tr ' =' '__')
  env $v python bench.py --resident-only --steps 4 > gpurun_out/r2e_bench_$f.json 2> gpurun_out/r2e_bench_$f.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2e_bench_$f.json").read().strip().splitlines()[-1])
print("VARIANT $v: frames/s %.0f render ms %.4f" % (d["value"], d["render_ms_per_launch"]))
P
done
