# Round-2 pass B (one B200): full GPU tests, non-integer pixel pitch at C4 scale, detector generator variants
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2b_tests.log; tail -3 gpurun_out/r2b_tests.log
python tools/microbench.py pitch > gpurun_out/r2b_pitch.jsonl 2> gpurun_out/r2b_pitch.err; cat gpurun_out/r2b_pitch.jsonl | cut -c1-700; tail -c 600 gpurun_out/r2b_pitch.err
for r in 10 7 1; do SCB_DETECTOR_ROUNDS=$r python tools/microbench.py detector_block; done > gpurun_out/r2b_detector_rounds.jsonl 2>&1
cat gpurun_out/r2b_detector_rounds.jsonl | cut -c1-400
SCB_DETECTOR_ROUNDS=7 python -m pytest tests/test_gpu_detector.py -x -q 2>&1 | tail -3 > gpurun_out/r2b_tests_detector_rounds7.log; cat gpurun_out/r2b_tests_detector_rounds7.log
