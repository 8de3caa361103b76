set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_forward_tc_r1d python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fwdtc.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/ncu_fwdtc.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -s -k "module_forwards or multivariate" 2>&1 | tail -8
