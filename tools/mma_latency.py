"""Latency of ONE short tcgen05.mma batch: first issue -> tcgen05.commit -> mbarrier wake-up of the issuing thread
(probe kernel, clock64).  The cross-attention kernels are chains of such batches (S = QK^T -> softmax -> PV -> store), so
this fixed cost, not the tensor throughput, is what bounds them."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import mma_cost as M  # noqa: E402  (prints its own table first)

print("---- latency of a single batch (reps = 1): cycles from first issue to wake-up ----")
for name, cfg in [
    ("SS N=80 K=48 (xattn QK^T d=40: 3 MMAs)", ((128, 40), "k", (80, 40), "k", 80, 48)),
    ("SS N=80 K=80 (d=80: 5 MMAs)", ((128, 80), "k", (80, 80), "k", 80, 80)),
    ("SS N=128 K=16 (1 MMA)", ((128, 16), "k", (128, 16), "k", 128, 16)),
    ("TS N=48 K=80 (xattn PV d=40: 5 MMAs)", ((128, 80), "tmem", (80, 40), "mn", 48, 80)),
    ("TS N=80 K=80 (d=80: 5 MMAs)", ((128, 80), "tmem", (80, 80), "mn", 80, 80)),
]:
    nk = cfg[5] // 16
    for reps in (1, 2, 8):
        per = M.cost(*cfg, reps=reps)
        print(f"{name:44s} reps={reps}: {per * reps * nk:8.0f} cycles total")
