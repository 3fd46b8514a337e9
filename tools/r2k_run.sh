set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "longest or chr19 or sweep_matches_oracle" > gpurun_out/r2k_tests.txt 2>&1
SWEEP_M=50 SWEEP_N=9958257 SWEEP_REPS=2 python tools/nsub_sweep.py > gpurun_out/r2k_probe_m50_chr1.txt 2>&1
SWEEP_M=200 SWEEP_N=2344705 SWEEP_REPS=2 python tools/nsub_sweep.py > gpurun_out/r2k_probe_m200.txt 2>&1
SWEEP_M=1000 SWEEP_N=500000 SWEEP_REPS=2 python tools/nsub_sweep.py > gpurun_out/r2k_probe_m1000.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2k_split_2gpu.json 2> gpurun_out/r2k_split_2gpu.err
python tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2k_split_1gpu.json 2> gpurun_out/r2k_split_1gpu.err
