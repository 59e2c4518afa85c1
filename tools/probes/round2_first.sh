# Round-2 first GPU pass: tests, weak bench, strong bench, C5 (one B200)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests.log
python bench.py > gpurun_out/r2_bench_weak.json 2> gpurun_out/r2_bench_weak.err
python bench.py --scaling strong --movie-frames 10000 > gpurun_out/r2_bench_strong.json 2> gpurun_out/r2_bench_strong.err
python bench.py --workload C5 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
tail -5 gpurun_out/r2_tests.log; tail -c 1500 gpurun_out/r2_bench_weak.err; tail -c 600 gpurun_out/r2_bench_strong.err; tail -c 600 gpurun_out/r2_bench_c5.err
for r in 10 7 1; do SCB_DETECTOR_ROUNDS=$r python tools/microbench.py detector_block; done > gpurun_out/r2_detector_rounds.jsonl 2>&1
SCB_DETECTOR_ROUNDS=7 python -m pytest tests/test_gpu_detector.py -x -q 2>&1 | tail -5 > gpurun_out/r2_tests_detector_rounds7.log
cat gpurun_out/r2_detector_rounds.jsonl gpurun_out/r2_tests_detector_rounds7.log
