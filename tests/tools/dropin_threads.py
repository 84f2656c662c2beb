#!/usr/bin/env python
"""GPU-box tool: wall clock of the drop-in program against the number of worker threads of the
reference (-tN): the workers block on their DP calls, so more of them = more problems in flight on
the device.  usage: dropin_threads.py [-A2|-A0] [n_cdna] [threads ...]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import realdata  # noqa: E402

alg = sys.argv[1] if len(sys.argv) > 1 else "-A2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
threads = [int(x) for x in sys.argv[3:]] or [16, 64, 256]
w = realdata.Workspace()
q = w.head_fasta(realdata.SEQDB / "dictdisc.cf", n)
base = None
for binary, tl in (("spaln", [w.threads]), ("spaln_gpu", threads)):
    for t in tl:
        opts = ["-Q7", "-O4", "-S3", alg, f"-t{t}", "-pq", "-Tdictdisc"]
        st = {}
        t0 = time.time()
        out = w.run(binary, opts, q, stats=st if binary == "spaln_gpu" else None)
        dt = time.time() - t0
        key = sorted(out.splitlines())
        if base is None:
            base = key
        print(f"{binary} {alg} -t{t}: {dt:.1f} s, output {'identical' if key == base else 'DIFFERS'}",
              st.get("set-up", ""), flush=True)
w.close()
