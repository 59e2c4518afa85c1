# four B200s, weak line only (sanity of the planned binning under torchrun)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2n4_bench_weak.json 2> gpurun_out/r2n4_bench_weak.err
tail -c 300 gpurun_out/r2n4_bench_weak.err
