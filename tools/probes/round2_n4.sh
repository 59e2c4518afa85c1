# Eight B200s: weak (the driver's command) and strong bench lines (torchrun, one rank per GPU)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2n4_bench_weak.json 2> gpurun_out/r2n4_bench_weak.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --scaling strong --movie-frames 10000 > gpurun_out/r2n4_bench_strong.json 2> gpurun_out/r2n4_bench_strong.err
python - <<'P'
import json
for f in ("weak", "strong"):
    try:
        d = json.loads(open("gpurun_out/r2n4_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], "e2e", json.dumps(d["e2e"])[:300], "gather_ok", d.get("gather_ok"), "export", json.dumps({k: v["value"] for k, v in d.get("export", {}).items()}), "host", json.dumps(d.get("host"))[:400])
    except Exception as e:
        print(f, "failed", e)
P
tail -c 500 gpurun_out/r2n4_bench_weak.err gpurun_out/r2n4_bench_strong.err
