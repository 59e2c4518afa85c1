set -x
for f in 8 16 32 64; do
  SCB_FRAMES_PER_LAUNCH=$f timeout 300 python bench.py --resident-only --steps 4 > gpurun_out/r2k_bench_$f.json 2> gpurun_out/r2k_bench_$f.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2k_bench_$f.json").read().strip().splitlines()[-1])
print("FPL $f: frames/s %.0f render ms/launch %.4f step ms %.3f" % (d["value"], d["render_ms_per_launch"], d["ms_per_step"]))
P
done
