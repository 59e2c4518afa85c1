set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_detector.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -x -q 2>&1 | tail -8
timeout 300 python tools/microbench.py detector > gpurun_out/r2i_detector.jsonl 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2i_detector.jsonl'):
    try: d=json.loads(l)
    except: continue
    if 'signal' in d and d['fpn']!='column': print(d['kernel'], d['size'], d['signal'], round(d['ms'],4), 'ms')
P
