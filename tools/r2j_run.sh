set -x
python -m pytest tests/test_sharding_gpu.py tests/test_gpu_parity.py tests/test_lean_sweeps.py -x -q > gpurun_out/r2j_tests.txt 2>&1
python tools/nsub_sweep.py > gpurun_out/r2j_probe_m10.txt 2>&1
SWEEP_M=50 SWEEP_N=9958257 SWEEP_REPS=2 python tools/nsub_sweep.py > gpurun_out/r2j_probe_m50_chr1.txt 2>&1
SWEEP_M=200 SWEEP_N=2344705 SWEEP_REPS=2 python tools/nsub_sweep.py > gpurun_out/r2j_probe_m200.txt 2>&1
SWEEP_M=1000 SWEEP_N=500000 SWEEP_REPS=2 python tools/nsub_sweep.py > gpurun_out/r2j_probe_m1000.txt 2>&1
