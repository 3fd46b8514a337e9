set -x
python -m pytest tests/test_sharding_gpu.py tests/test_background_gpu.py -x -q > gpurun_out/r2v_tests.txt 2>&1
python tools/bg_singular_probe.py > gpurun_out/r2v_bg_singular.txt 2>&1
python tools/split_ecm_nccl.py chr21 500 4 > gpurun_out/r2v_split_tiny.json 2> gpurun_out/r2v_split_tiny.err
python - > gpurun_out/r2v_bg_time.txt 2>&1 <<'PY'
import time, numpy as np, consenrich_b200 as cb
rng=np.random.default_rng(0); n=2344705
w=rng.uniform(0.5,2,n); r=rng.normal(size=n)
for _ in range(3): cb.csolveZeroCenteredBackground(w,r,128.0,True)
t0=time.perf_counter()
for _ in range(5): cb.csolveZeroCenteredBackground(w,r,128.0,True)
print("host-api solve ms", (time.perf_counter()-t0)/5*1e3)
PY
