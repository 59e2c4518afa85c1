# occupancy bounds of the binning kernels: rebuild render.cu with -D flags on the box, time the resident bench
set -x
for cfg in "2 1" "4 1" "3 1" "4 6"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -O3 --expt-relaxed-constexpr -DSCB_FILL_CTAS=$1 -DSCB_PREPARE_CTAS=$2 -c scopyon_b200/csrc/render.cu -o scopyon_b200/_build/render.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scopyon_b200/libscopyon_b200.so scopyon_b200/_build/*.o -lcuda
  timeout 300 python bench.py --resident-only --steps 6 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
  python - <<P
import json
d=json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
print("OCC fill $1 prepare $2: frames/s %.0f render ms/launch %.4f step ms %.3f rest ms/launch %.4f" % (d["value"], d["render_ms_per_launch"], d["ms_per_step"], d["ms_per_step"]/4 - d["render_ms_per_launch"]))
P
done
