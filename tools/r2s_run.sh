set -x
python -m pytest tests/test_gpu_parity.py tests/test_munc_gpu.py -x -q > gpurun_out/r2s_tests.txt 2>&1
for cfg in "10 2344705" "50 9958257" "200 2344705" "1000 500000" "7 1000003" "33 500001"; do set -- $cfg; SWEEP_M=$1 SWEEP_N=$2 SWEEP_REPS=2 python tools/nsub_sweep.py >> gpurun_out/r2s_probe.txt 2>&1; done
bash tools/r2_configs.sh
