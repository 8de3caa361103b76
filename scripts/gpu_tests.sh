# usage: bash scripts/gpu_tests.sh [pytest -k expression]
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -rA ${1:+-k "$1"} > gpurun_out/pytest_gpu_full.log 2>&1; grep -v "^PASSED" gpurun_out/pytest_gpu_full.log | tail -120 > gpurun_out/pytest_gpu.log; tail -60 gpurun_out/pytest_gpu.log
