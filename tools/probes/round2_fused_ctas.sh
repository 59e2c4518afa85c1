# CTAs per SM of the fused binning kernel: resident bench per variant
for v in 4 3 5; do
  SCB_FUSED_CTAS=$v timeout 300 python bench.py --resident-only --steps 6 > gpurun_out/r2v_bench_$v.json 2> gpurun_out/r2v_bench_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2v_bench_$v.json").read().strip().splitlines()[-1])
print("FUSED_CTAS $v: frames/s %.0f step ms %.3f render %.4f" % (d["value"], d["ms_per_step"], d["render_ms_per_launch"]))
P
done
