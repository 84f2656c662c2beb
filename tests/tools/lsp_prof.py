import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, bench
from spaln_b200 import Engine
prm = bench.load_params()
raw = bench.make_workload(10000, 20251017); bench.host_cells(raw)
P = bench.to_problems(raw)
eng = Engine(prm, 0)
eng.lspS_ng(P[:200])
for _ in range(2):
    t0=time.perf_counter(); r = eng.lspS_ng(P, max_vmf_space=32*1024*1024, sh=100, alg=2); print("python total ms", 1e3*(time.perf_counter()-t0))
