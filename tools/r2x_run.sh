set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2x_gpu_tests.txt 2>&1
python tools/split_ecm_nccl.py chr21 500 4 > gpurun_out/r2x_split_tiny.json 2> gpurun_out/r2x_split_tiny.err
SPLIT_GRAPHS=0 python tools/split_ecm_nccl.py chr21 500 4 > gpurun_out/r2x_split_tiny_nograph.json 2> gpurun_out/r2x_split_tiny_nograph.err
python tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2x_split_1gpu.json 2> gpurun_out/r2x_split_1gpu.err
