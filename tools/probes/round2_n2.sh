# Two B200s: NCCL gather test on hardware, weak and strong bench lines (torchrun, one rank per GPU)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_movie.py -q -k "gather_frames_nccl" 2>&1 | tail -3 > gpurun_out/r2n2_gather_test.log; cat gpurun_out/r2n2_gather_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2n2_bench_weak.json 2> gpurun_out/r2n2_bench_weak.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --scaling strong --movie-frames 10000 > gpurun_out/r2n2_bench_strong.json 2> gpurun_out/r2n2_bench_strong.err
python - <<'P'
import json
for f in ("weak", "strong"):
    try:
        d = json.loads(open("gpurun_out/r2n2_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], "e2e", d["e2e"]["value"], "gather_ok", d.get("gather_ok"), "export", json.dumps(d.get("export"))[:400])
    except Exception as e:
        print(f, "failed", e)
P
tail -c 600 gpurun_out/r2n2_bench_weak.err gpurun_out/r2n2_bench_strong.err
