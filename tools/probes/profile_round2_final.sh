# Round-2 final measurement pass (after the shared-memory census and the up-front fill) on one B200; outputs under gpurun_out/r2g_*, summaries go to profiles/
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2g_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2g_smoke.log 2>&1; tail -2 gpurun_out/r2g_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_weak.json 2> gpurun_out/r2g_bench_weak.err
python bench.py --scaling strong --movie-frames 10000 > gpurun_out/r2g_bench_strong.json 2> gpurun_out/r2g_bench_strong.err
python bench.py --workload C5 > gpurun_out/r2g_bench_c5.json 2> gpurun_out/r2g_bench_c5.err
python bench.py --impl reference > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err
# launch list (cold, serialised): 16-frame block launches, then the API path (blocks of 8 frames)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 1 --warmup 1 --frames-per-step 16 --e2e-frames 8 --no-cpu-baseline > gpurun_out/r2g_launches.log 2>&1
# full metrics of the four main kernels of the second 16-frame block
ncu --set full --clock-control none --import-source on -k regex:'render_strips|detector_fast|strip_fill|spot_prepare|spot_bin_fused' -s 0 -c 8 -f -o gpurun_out/r2g_block python bench.py --resident-only --steps 2 --warmup 1 --frames-per-step 16 > gpurun_out/r2g_block.log 2>&1
timeout 600 python tools/microbench.py detector diffuse render pitch > gpurun_out/r2g_microbench.jsonl 2>&1
python tools/config_timings.py > gpurun_out/r2g_config_timings.jsonl 2>&1
cat gpurun_out/r2g_tests.log gpurun_out/r2g_config_timings.jsonl
