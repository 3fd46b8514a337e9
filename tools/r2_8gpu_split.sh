set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29531 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2u_split_8gpu.json 2> gpurun_out/r2u_split_8gpu.err
$TR --nproc-per-node 8 --master-port 29532 tools/split_ecm_nccl.py chr1 10 50 > gpurun_out/r2u_split_8gpu_m50.json 2> gpurun_out/r2u_split_8gpu_m50.err
$TR --nproc-per-node 4 --master-port 29533 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2u_split_4gpu.json 2> gpurun_out/r2u_split_4gpu.err
$TR --nproc-per-node 2 --master-port 29534 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2u_split_2gpu.json 2> gpurun_out/r2u_split_2gpu.err
