set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2w_gpu_tests.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.txt 2>&1
python bench.py > gpurun_out/r2w_bench_cfg3.json 2> gpurun_out/r2w_bench_cfg3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err
