"""tcgen05.mma cost per instruction for the operand flavours the attention kernels use (probe kernel, clock64)."""
import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from diffusion_spacetime_attn_b200 import native, umma_enc as E
import test_probe_gpu as T

def cost(a_shape, a_mode, b_shape, b_mode, n, k_mma, reps=64, alt=False):
    A, B = T._rand(a_shape, 1), T._rand(b_shape, 2)
    lib = native.load(); args = native.ProbeArgs(); nk = k_mma // 16
    args.a, args.b = A.data_ptr(), B.data_ptr()
    args.a_rows = args.a_tensor_rows = A.shape[0]; args.a_cols = A.shape[1]; args.a_in_tmem = int(a_mode == "tmem")
    args.b_rows = args.b_tensor_rows = B.shape[0]; args.b_cols = B.shape[1]
    args.a_desc_hi = E.desc_hi_sw128(16, 1024) if a_mode == "k" else E.desc_hi_sw128(A.shape[0] * 128, 1024)
    a_off = E.kmajor_offsets(k_mma, A.shape[0]) if a_mode == "k" else (E.mnmajor_offsets(k_mma) if a_mode == "mn" else [8 * i for i in range(nk)])
    args.b_desc_hi = E.desc_hi_sw128(16, 1024) if b_mode == "k" else E.desc_hi_sw128(B.shape[0] * 128, 1024)
    b_off = E.kmajor_offsets(k_mma, B.shape[0]) if b_mode == "k" else E.mnmajor_offsets(k_mma)
    args.nk = nk
    for i in range(nk): args.a_off[i] = a_off[i]; args.b_off[i] = b_off[i]
    args.idesc = E.idesc_f16(128, n, int(a_mode == "mn"), int(b_mode == "mn")); args.n = n
    out = torch.zeros(128, n, device="cuda"); args.out = out.data_ptr()
    args.dump_bytes = -1 if alt else 0
    cyc = torch.zeros(1, dtype=torch.int64, device="cuda"); args.cycles = cyc.data_ptr(); args.reps = reps
    for _ in range(2):
        native.check(lib.sta_probe_gemm(C.byref(args), None), "probe"); torch.cuda.synchronize()
    return cyc.item() / (reps * nk)

for name, cfg in [
    ("SS K/K   N=128 (fwd QK^T d=40)", ((128, 40), "k", (128, 40), "k", 128, 48)),
    ("SS K/K   N=64  (bwd S^T half)", ((128, 40), "k", (64, 40), "k", 64, 48)),
    ("SS K/K   N=128 K=80", ((128, 80), "k", (128, 80), "k", 128, 80)),
    ("SS K/K   N=256 ", ((128, 64), "k", (256, 64), "k", 256, 64)),
    ("SS K/K   N=80  (xattn)", ((128, 40), "k", (80, 40), "k", 80, 48)),
    ("TS      N=48 MN-B (fwd PV d=40)", ((128, 128), "tmem", (128, 40), "mn", 48, 128)),
    ("TS      N=64 MN-B", ((128, 128), "tmem", (128, 64), "mn", 64, 128)),
    ("TS      N=80 MN-B (d=80)", ((128, 128), "tmem", (128, 80), "mn", 80, 128)),
    ("TS      N=128 MN-B", ((128, 128), "tmem", (128, 128), "mn", 128, 128)),
    ("SS K-A  N=48 MN-B (bwd dK)", ((128, 128), "k", (128, 40), "mn", 48, 128)),
    ("SS MN-A N=48 MN-B (bwd dQ)", ((128, 128), "mn", (128, 40), "mn", 48, 128)),
    ("SS MN-A N=128 MN-B", ((128, 128), "mn", (128, 128), "mn", 128, 128)),
]:
    print(f"{name:36s} {cost(*cfg):7.1f} cycles / MMA (K=16)   two alternating accumulators: {cost(*cfg, alt=True) if cfg[4] <= 128 else float('nan'):7.1f}")
