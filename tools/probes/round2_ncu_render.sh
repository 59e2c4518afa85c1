# ncu --set full of one 16-frame render launch per variant in $VARIANTS ("name:ENV=..,ENV=.." separated by spaces)
set -x
mkdir -p gpurun_out
for spec in $VARIANTS; do
  name=${spec%%:*}; envs=$(echo ${spec#*:} | tr ',' ' ')
  env $envs ncu --set full --clock-control none --import-source on -k regex:'render_strips|render_tiles' -s 2 -c 1 -f -o gpurun_out/r2_render_$name python bench.py --resident-only --steps 2 --warmup 1 --frames-per-step 16 > gpurun_out/r2_ncu_$name.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
