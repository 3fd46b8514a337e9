set -x
python -m pytest tests/test_sharding_gpu.py tests/test_lean_sweeps.py -x -q > gpurun_out/r2t_tests.txt 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -k "longest or chr19 or sweep_matches_oracle" >> gpurun_out/r2t_tests.txt 2>&1
for cfg in "10 2344705" "50 9958257" "200 2344705" "1000 500000" "150 700001"; do set -- $cfg; SWEEP_M=$1 SWEEP_N=$2 SWEEP_REPS=2 python tools/nsub_sweep.py >> gpurun_out/r2t_probe.txt 2>&1; done
python tools/bg_singular_probe.py > gpurun_out/r2t_bg_singular.txt 2>&1
python tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2t_split_1gpu.json 2> gpurun_out/r2t_split_1gpu.err
