"""GPU: pin the tcgen05 / TMA operand encodings of csrc/sta_common.cuh with single-tile GEMMs.

Each case stages fp16 operands exactly the way the attention kernels do (TMA 128B-swizzled boxes, zero-filled
out-of-bounds rows/columns, A optionally packed into TMEM) and compares the fp32 accumulator with torch.
Run standalone (`python tests/test_probe_gpu.py`) to get a table of all cases without stopping at a failure.
"""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import native, umma_enc as E  # noqa: E402


def _ceil(x, m):
    return (x + m - 1) // m * m


def run_probe(A, B, *, a_mode, b_mode, n, k_mma, a_box_rows=None, b_box_rows=None, dump=False):
    """A, B: fp16 cuda tensors as stored in global memory (see modes).  Returns (D[128,n] fp32, smem image|None).

    a_mode: 'k'    A is [128, K]    K-major, from shared memory
            'mn'   A is [K, 128]    M-major (transposed storage), from shared memory
            'tmem' A is [128, K]    packed into TMEM by the probe
    b_mode: 'k'    B is [N, K]      K-major   (D = A @ B^T)
            'mn'   B is [K, N]      N-major   (D = A @ B)
    """
    lib = native.load()
    args = native.ProbeArgs()
    nk = k_mma // 16
    assert nk <= 16
    a_rows = a_box_rows or A.shape[0]
    b_rows = b_box_rows or B.shape[0]
    args.a = A.data_ptr()
    args.a_rows, args.a_tensor_rows, args.a_cols = a_rows, A.shape[0], A.shape[1]
    args.a_in_tmem = 1 if a_mode == "tmem" else 0
    args.b = B.data_ptr()
    args.b_rows, args.b_tensor_rows, args.b_cols = b_rows, B.shape[0], B.shape[1]
    if a_mode == "k":
        args.a_desc_hi = E.desc_hi_sw128(16, 1024)
        a_off = E.kmajor_offsets(k_mma, a_rows)
    elif a_mode == "mn":
        args.a_desc_hi = E.desc_hi_sw128(a_rows * 128, 1024)
        a_off = E.mnmajor_offsets(k_mma)
    else:
        args.a_desc_hi = 0
        a_off = [8 * i for i in range(nk)]
    if b_mode == "k":
        args.b_desc_hi = E.desc_hi_sw128(16, 1024)
        b_off = E.kmajor_offsets(k_mma, b_rows)
    else:
        args.b_desc_hi = E.desc_hi_sw128(b_rows * 128, 1024)
        b_off = E.mnmajor_offsets(k_mma)
    args.nk = nk
    for i in range(nk):
        args.a_off[i] = a_off[i]
        args.b_off[i] = b_off[i]
    args.idesc = E.idesc_f16(128, n, 1 if a_mode == "mn" else 0, 1 if b_mode == "mn" else 0)
    args.n = n
    out = torch.full((128, n), float("nan"), device="cuda", dtype=torch.float32)
    args.out = out.data_ptr()
    img = None
    if dump:
        img = torch.zeros(128 * 1024, device="cuda", dtype=torch.uint8)
        args.smem_dump = img.data_ptr()
        args.dump_bytes = img.numel()
    native.check(lib.sta_probe_gemm(C.byref(args), None), "sta_probe_gemm")
    torch.cuda.synchronize()
    err = native.device_error()
    assert err == 0, f"device error word 0x{err:x}"
    return out, img


def _rand(shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(shape, generator=g).half().cuda()


def _expect(A, B, a_mode, b_mode, n):
    Af = A.float().t() if a_mode == "mn" else A.float()
    Bf = B.float() if b_mode == "mn" else B.float().t()
    D = Af @ Bf  # [128, n_real]
    if D.shape[1] < n:
        D = torch.nn.functional.pad(D, (0, n - D.shape[1]))
    return D[:, :n]


# name, a_shape, a_mode, b_shape, b_mode, n, k_mma, a_box_rows, b_box_rows
CASES = [
    ("qk_d40", (128, 40), "k", (128, 40), "k", 128, 48, None, None),
    ("qk_d80", (128, 80), "k", (128, 80), "k", 128, 80, None, None),
    ("qk_d160", (128, 160), "k", (128, 160), "k", 128, 160, None, None),
    ("qk_ctx77_d40", (128, 40), "k", (77, 40), "k", 80, 48, None, 80),
    ("qk_ctx77_d160", (128, 160), "k", (77, 160), "k", 80, 160, None, 80),
    ("qk_rows_oob", (100, 80), "k", (128, 80), "k", 128, 80, 128, None),
    ("pv_ss_d40_n48", (128, 128), "k", (128, 40), "mn", 48, 128, None, None),
    ("pv_ss_d40_n64", (128, 128), "k", (128, 40), "mn", 64, 128, None, None),
    ("pv_ss_d80_n80", (128, 128), "k", (128, 80), "mn", 80, 128, None, None),
    ("pv_ss_d80_n128", (128, 128), "k", (128, 80), "mn", 128, 128, None, None),
    ("pv_ss_d160_n160", (128, 128), "k", (128, 160), "mn", 160, 128, None, None),
    ("pv_ts_d40_n48", (128, 128), "tmem", (128, 40), "mn", 48, 128, None, None),
    ("pv_ts_d80_n80", (128, 128), "tmem", (128, 80), "mn", 80, 128, None, None),
    ("pv_ts_d160_n160", (128, 128), "tmem", (128, 160), "mn", 160, 128, None, None),
    ("pv_ts_ctx77_d40", (128, 80), "tmem", (77, 40), "mn", 48, 80, None, 80),
    ("pv_ts_ctx77_d160", (128, 80), "tmem", (77, 160), "mn", 160, 80, None, 80),
    ("dq_mmajor_d40", (128, 128), "mn", (128, 40), "mn", 48, 128, None, None),
    ("dq_mmajor_d160", (128, 128), "mn", (128, 160), "mn", 160, 128, None, None),
    ("dq_mmajor_ctx80", (80, 128), "mn", (77, 80), "mn", 80, 80, None, 80),
    ("dkv_kmajor_mn_d80", (128, 128), "k", (128, 80), "mn", 80, 128, None, None),
]


def _run_case(case):
    name, a_shape, a_mode, b_shape, b_mode, n, k_mma, a_box, b_box = case
    A = _rand(a_shape, 1)
    B = _rand(b_shape, 2)
    if a_mode == "tmem" and a_shape[1] == 80:
        A[:, 77:] = 0  # what the kernels store for padded keys
    D, _ = run_probe(A, B, a_mode=a_mode, b_mode=b_mode, n=n, k_mma=k_mma, a_box_rows=a_box, b_box_rows=b_box)
    Aeff, Beff = A, B
    if b_box and b_box > B.shape[0]:  # TMA zero-fills the out-of-bounds rows of the box
        Beff = torch.nn.functional.pad(B, (0, 0, 0, b_box - B.shape[0]))
    ref = _expect(Aeff, Beff, a_mode, b_mode, n)
    if ref.shape[0] < 128:
        ref = torch.nn.functional.pad(ref, (0, 0, 0, 128 - ref.shape[0]))
    kk = min(Aeff.shape[1] if a_mode != "mn" else Aeff.shape[0], 10_000)
    tol = 2e-2 * (kk ** 0.5)
    err = (D - ref).abs().max().item()
    return err, tol


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_probe_gemm(case):
    err, tol = _run_case(case)
    assert err == err and err < 1e-2, f"max abs err {err}"


@pytest.mark.gpu
def test_tma_swizzle_image():
    """The bytes TMA leaves in shared memory follow sw128_offset(row, chunk), block after block."""
    A = _rand((128, 80), 3)
    B = _rand((128, 80), 4)
    _, img = run_probe(A, B, a_mode="k", b_mode="k", n=128, k_mma=80, dump=True)
    img = img.cpu()
    a_bytes = A.cpu().contiguous().view(torch.uint8).reshape(128, 160)
    want = torch.zeros(2 * 128 * 128, dtype=torch.uint8)
    for blk in range(2):
        for r in range(128):
            for c in range(8):
                col0 = blk * 128 + c * 16
                chunk = torch.zeros(16, dtype=torch.uint8)
                if col0 < 160:
                    chunk = a_bytes[r, col0:col0 + 16]
                off = blk * 128 * 128 + E.sw128_offset(r, c)
                want[off:off + 16] = chunk
    assert torch.equal(img[: want.numel()], want)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for case in CASES:
        try:
            err, tol = _run_case(case)
            print(f"{case[0]:24s} max_abs_err={err:.4e}  {'OK' if err < 1e-2 else 'MISMATCH'}", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"{case[0]:24s} EXC {ex}", flush=True)
    try:
        test_tma_swizzle_image()
        print("tma_swizzle_image        OK")
    except Exception as ex:  # noqa: BLE001
        print("tma_swizzle_image        FAIL", repr(ex)[:200])
