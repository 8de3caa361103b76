# usage: bash scripts/gpu_cfg_timings.sh TAG -- wall-clock of configs 1/2/4/5 on one GPU + the launch list of a config-1 signal
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 600 python scripts/config_timings.py > gpurun_out/config_timings_$TAG.json 2> gpurun_out/config_timings_$TAG.err; cat gpurun_out/config_timings_$TAG.json; tail -3 gpurun_out/config_timings_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg1_$TAG.csv python scripts/cfg1_once.py 6 > gpurun_out/ncu_cfg1.log 2>&1; echo "ncu exit $?"
