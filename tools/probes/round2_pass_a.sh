# Round-2 pass A on one B200: GPU tests, weak / strong / C5 bench lines, reference arm, launch list.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_weak.json 2> gpurun_out/r2a_bench_weak.err
python bench.py --scaling strong --movie-frames 10000 > gpurun_out/r2a_bench_strong.json 2> gpurun_out/r2a_bench_strong.err
python bench.py --workload C5 > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err
python bench.py --half-life 2.5 --resident-only > gpurun_out/r2a_bench_halflife.json 2> gpurun_out/r2a_bench_halflife.err
python bench.py --impl reference > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 1 --warmup 1 --frames-per-step 16 --e2e-frames 4 --no-cpu-baseline > gpurun_out/r2a_launches.log 2>&1
tail -5 gpurun_out/r2a_tests.log
for f in weak strong c5 halflife reference; do echo "== $f"; cut -c1-3000 gpurun_out/r2a_bench_$f.json; tail -c 800 gpurun_out/r2a_bench_$f.err; done
