# Profiling pass of round 2 (run on the GPU box through gpurun; reports are summarised there, only the
# summaries and the launch list come back).
set -x
B="--steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
python bench.py $B > gpurun_out/r2p_bench_profiled_command.json 2> gpurun_out/r2p_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fold_kernel|lean_|residual_kernel|scan_kernel" --csv \
    --log-file gpurun_out/r2p_launches.csv python bench.py $B > gpurun_out/r2p_ncu_bench.log 2>&1
export SWEEP_M=50 SWEEP_N=9958257
python tools/ecm_once.py 2 > gpurun_out/r2p_once.txt 2>&1 || exit 1
ncu --set full --clock-control none -k regex:"fold_kernel|lean_|residual_kernel" --launch-skip 83 --launch-count 10 \
    -o /tmp/r2p_a python tools/ecm_once.py 2 > gpurun_out/r2p_ncu_a.log 2>&1
ncu --set full --clock-control none -k regex:"residual_kernel|lean_bwd_replay_kernel" --launch-skip 32 --launch-count 2 \
    -o /tmp/r2p_b python tools/ecm_once.py 2 > gpurun_out/r2p_ncu_b.log 2>&1
python tools/ncu_summary2.py gpurun_out/r2_ncu_full_summary.json 50 9958257 \
    "ncu --set full --clock-control none of the second of two cb200_ecm_device calls (tools/ecm_once.py; 50 tracks x chr1 @ 25 bp, K=3, t=5): its first ten kernels and its last two (tools/r2_profile.sh)" \
    /tmp/r2p_a.ncu-rep /tmp/r2p_b.ncu-rep > gpurun_out/r2p_summary.log 2>&1
ls -la /tmp/*.ncu-rep >> gpurun_out/r2p_summary.log
export SWEEP_M=200 SWEEP_N=2344705
ncu --set full --clock-control none -k regex:"residual_kernel|fold_kernel" --launch-skip 2 --launch-count 2 \
    -o /tmp/r2p_c python tools/ecm_once.py 2 > gpurun_out/r2p_ncu_c.log 2>&1
python tools/ncu_summary2.py gpurun_out/r2_ncu_m200_summary.json 200 2344705 \
    "ncu --set full --clock-control none: fold and residual kernels of the second of two cb200_ecm_device calls, 200 tracks x chr19 @ 25 bp" \
    /tmp/r2p_c.ncu-rep >> gpurun_out/r2p_summary.log 2>&1
