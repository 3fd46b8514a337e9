# 8 GPUs of one box: the whole-genome benchmark strong-scaled, and chr1 @ 10 bp split into 8 ranges
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 > gpurun_out/r2r_bench_8gpu.json 2> gpurun_out/r2r_bench_8gpu.err
$TR --nproc-per-node 8 --master-port 29522 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2r_split_8gpu.json 2> gpurun_out/r2r_split_8gpu.err
$TR --nproc-per-node 4 --master-port 29523 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2r_split_4gpu.json 2> gpurun_out/r2r_split_4gpu.err
$TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 5 > gpurun_out/r2r_bench_4gpu.json 2> gpurun_out/r2r_bench_4gpu.err
