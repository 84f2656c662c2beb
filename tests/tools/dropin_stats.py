#!/usr/bin/env python
"""GPU-box tool: run the drop-in program (oracle/_ref/spaln_gpu) on the sample data and print how
many calls each hook answered on the device.  usage: dropin_stats.py [-A0|-A1|-A2|-A3] [protein|cdna] [n_cdna]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import realdata  # noqa: E402

alg = sys.argv[1] if len(sys.argv) > 1 else "-A0"
what = sys.argv[2] if len(sys.argv) > 2 else "protein"
w = realdata.Workspace()
if what == "protein":
    q, opts = realdata.SEQDB / "dictdisc.faa", ["-Q7", "-O0", alg, "-t1", "-pq", "-Tdictdisc"]
else:
    q = w.head_fasta(realdata.SEQDB / "dictdisc.cf", int(sys.argv[3]) if len(sys.argv) > 3 else 200)
    opts = ["-Q7", "-O4", "-S3", alg, f"-t{w.threads}", "-pq", "-Tdictdisc"]
for binary in ("spaln", "spaln_gpu"):
    st = {}
    t0 = time.time()
    out = w.run(binary, opts, q, stats=st)
    print(binary, " ".join(opts), f"{time.time() - t0:.1f} s", len(out), "bytes", st, flush=True)
    if st.get("set-up"):
        print("   set-up:", st["set-up"], flush=True)
w.close()
