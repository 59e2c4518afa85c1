# Round-end measurement pass on one B200 (run through gpurun); outputs under gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/tests_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
# launch list (cold, serialised): the first two blocks are the 16-frame launches, the rest the per-frame API path
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --frames-per-step 16 --e2e-frames 4 > gpurun_out/launches_final.log 2>&1
# full metrics of the four main kernels of the second 16-frame block
ncu --set full --clock-control none --import-source on -k regex:'render_strips|detector_fast|strip_fill|spot_prepare' -s 4 -c 4 -f -o gpurun_out/block_final python bench.py --steps 1 --warmup 1 --frames-per-step 16 --e2e-frames 2 > gpurun_out/block_final.log 2>&1
python tools/microbench.py > gpurun_out/microbench_final.jsonl 2>&1
python tools/config_timings.py > gpurun_out/config_timings_final.jsonl 2>&1
cat gpurun_out/tests_final.log gpurun_out/config_timings_final.jsonl
