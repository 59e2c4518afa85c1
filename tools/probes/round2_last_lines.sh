mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_weak.json 2> gpurun_out/r2h_bench_weak.err
python bench.py --scaling strong --movie-frames 10000 > gpurun_out/r2h_bench_strong.json 2> gpurun_out/r2h_bench_strong.err
timeout 300 python -m pytest tests/test_gpu_movie.py -x -q 2>&1 | tail -2
