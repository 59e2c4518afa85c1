set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_render.py tests/test_gpu_geometries.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4
timeout 300 python tools/microbench.py pitch > gpurun_out/r2j_pitch.jsonl 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2j_pitch.jsonl'):
    try: d=json.loads(l)
    except: continue
    if 'pixel' in d: print(d['pixel'], round(d['ms'],4), 'ms', d['checksum'])
P
timeout 300 python bench.py --resident-only --steps 4 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[-1])
print("C4: frames/s %.0f render ms %.4f step ms %.3f" % (d["value"], d["render_ms_per_launch"], d["ms_per_step"]))
P
