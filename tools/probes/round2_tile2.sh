set -x
mkdir -p gpurun_out
for v in "SCB_RENDER_PATH=tile"; do
  f=$(echo $v | tr ' =' '__')
  env $v timeout 300 python bench.py --resident-only --steps 4 > gpurun_out/r2g_bench_$f.json 2> gpurun_out/r2g_bench_$f.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2g_bench_$f.json").read().strip().splitlines()[-1])
print("VARIANT $v: frames/s %.0f render ms %.4f step ms %.3f" % (d["value"], d["render_ms_per_launch"], d["ms_per_step"]))
P
done
