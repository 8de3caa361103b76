# usage: bash scripts/gpu_ncu_kernel.sh <kernel regex> <tag>   -- one ncu --set full capture of the 4th launch of the kernel under bench.py
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -f -o gpurun_out/prof_$2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$2.log 2>&1; echo "ncu exit $?"
