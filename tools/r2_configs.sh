# cfg4 / cfg5 on one GPU (device-resident numbers; the e2e leg of these sizes needs > 100 GB of pinned host memory)
set -x
python bench.py --config cfg4 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2q_bench_cfg4.json 2> gpurun_out/r2q_bench_cfg4.err
python bench.py --config cfg5 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2q_bench_cfg5.json 2> gpurun_out/r2q_bench_cfg5.err
python bench.py --config cfg2 --steps 20 > gpurun_out/r2q_bench_cfg2.json 2> gpurun_out/r2q_bench_cfg2.err
