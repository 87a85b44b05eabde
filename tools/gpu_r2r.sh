#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_memory.py tests/test_train.py -m gpu -x -q 2>&1 | tail -15
timeout 120 python - <<'PY'
import torch, numpy as np, time, ctypes as C
from cumind_b200.memory import TreeBuffer
dev=torch.device("cuda",0)
for cap in (2000, 8192, 100000):
    b=TreeBuffer(cap,0.6,1e-6,device=dev)
    b.add_batch(list(range(min(cap,4096))))
    rng=np.random.default_rng(0)
    t=b._dev
    for n in (32, 1024, 4096):
        idx=torch.as_tensor(rng.integers(0,min(cap,4096),size=n)+b.tree_size-1).to(dev)
        pri=torch.as_tensor(rng.random(n)).to(dev)
        st=C.c_void_p(torch.cuda.current_stream().cuda_stream)
        call=lambda: t._lib.cumind_sumtree_update(t._handle, C.c_void_p(idx.data_ptr()), C.c_void_p(pri.data_ptr()), C.c_int64(n), st)
        for _ in range(3): call()
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): call()
        e1.record(); torch.cuda.synchronize()
        print(f"cap {cap} n {n}: {e0.elapsed_time(e1)/20*1e3:.1f} us per update launch")
PY
