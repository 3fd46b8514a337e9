set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29541 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2y_split_8gpu_graphs.json 2> gpurun_out/r2y_split_8gpu_graphs.err
SPLIT_GRAPHS=0 $TR --nproc-per-node 8 --master-port 29542 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2y_split_8gpu_eager.json 2> gpurun_out/r2y_split_8gpu_eager.err
$TR --nproc-per-node 4 --master-port 29543 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2y_split_4gpu_graphs.json 2> gpurun_out/r2y_split_4gpu_graphs.err
$TR --nproc-per-node 2 --master-port 29544 tools/split_ecm_nccl.py chr1 10 4 > gpurun_out/r2y_split_2gpu_graphs.json 2> gpurun_out/r2y_split_2gpu_graphs.err
