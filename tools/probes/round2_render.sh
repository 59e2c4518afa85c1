# Render variants: strip shape x copy engine (one B200)
set -x
python -m pytest tests/test_gpu_detector.py tests/test_gpu_api.py tests/test_gpu_movie.py tests/test_gpu_particles.py -q 2>&1 | tail -25 > gpurun_out/r2_tests_b.log
for rows in 8 16; do for copy in tma ldgsts; do
  export SCB_RENDER_ROWS=$rows SCB_RENDER_COPY=$copy
  python -m pytest tests/test_gpu_render.py tests/test_gpu_geometries.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4 > gpurun_out/r2_tests_render_${rows}_${copy}.log
  python bench.py --resident-only --steps 4 > gpurun_out/r2_bench_render_${rows}_${copy}.json 2> gpurun_out/r2_bench_render_${rows}_${copy}.err
done; done
unset SCB_RENDER_ROWS SCB_RENDER_COPY
tail -3 gpurun_out/r2_tests_b.log gpurun_out/r2_tests_render_*.log
cat gpurun_out/r2_bench_render_*.json
tail -c 400 gpurun_out/r2_bench_render_*.err
python bench.py > gpurun_out/r2_bench_weak_b.json 2> gpurun_out/r2_bench_weak_b.err; cut -c1-1500 gpurun_out/r2_bench_weak_b.json | grep -o '"e2e".*' | cut -c1-900; tail -c 600 gpurun_out/r2_bench_weak_b.err
