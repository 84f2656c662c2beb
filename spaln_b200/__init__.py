"""spaln_b200 -- B200-native spliced-alignment DP engine (gspaln).

Only what the DP hot path needs lives here:
  csrc/       hand-written sm_100a CUDA kernels + the C-ABI (include/gspaln.h)
  capi.py     ctypes binding of that C-ABI
  engine.py   host-side mirror of the reference's SimdAln2s1 / SimdAln2h1 call surface (Engine,
              EngineH), of the Exinon constructor (ExinonScan, ExinonScanP) and of Seq::nuc2tron
  workload.py seeded synthetic problems of the BASELINE.json shapes
"""
from .capi import FORWARD_WIP, SCOREONLY_WIP  # noqa: F401
from .engine import (Engine, EngineError, EngineH, ExinonScan, ExinonScanP, PackedBatch, Problem,  # noqa: F401
                     ProblemH, Result, Timing, nuc2tron)

__version__ = "0.1.0"
