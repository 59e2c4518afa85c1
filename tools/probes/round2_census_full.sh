# full ncu capture of spot_prepare (shared-memory census) and strip_fill on one 16-frame block
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'spot_prepare|strip_fill' -s 2 -c 2 -f -o gpurun_out/r2s_binning python bench.py --resident-only --steps 2 --warmup 1 --frames-per-step 16 > gpurun_out/r2s_binning.log 2>&1
