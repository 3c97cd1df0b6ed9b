"""Host-side mirror of the UMMA descriptor encodings in csrc/sta_common.cuh (used by the probe tests)."""
from __future__ import annotations


def desc_hi_sw128(lbo_bytes: int, sbo_bytes: int) -> int:
    """Shared-memory matrix descriptor template (start address = 0), 128B swizzle, sm_100 version bit."""
    return (((lbo_bytes >> 4) & 0x3FFF) << 16) | (((sbo_bytes >> 4) & 0x3FFF) << 32) | (1 << 46) | (2 << 61)


def idesc_f16(m: int, n: int, a_mn_major: int = 0, b_mn_major: int = 0) -> int:
    """Instruction descriptor, kind::f16, fp16 operands, fp32 accumulator."""
    return (1 << 4) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def kmajor_offsets(k_total: int, rows: int) -> list[int]:
    """Byte offset of each K=16 step for a K-major operand staged as 64-column blocks of `rows` rows."""
    offs = []
    for k in range(0, k_total, 16):
        blk, within = divmod(k, 64)
        offs.append(blk * rows * 128 + within * 2)
    return offs


def mnmajor_offsets(k_total: int) -> list[int]:
    """Byte offset of each K=16 step for an MN-major operand (rows = K index): two 8-row groups per step."""
    return [(k // 8) * 1024 for k in range(0, k_total, 16)]


def sw128_offset(r: int, c: int) -> int:
    """Byte offset of (row r, 16-byte chunk c) inside one swizzled block."""
    return r * 128 + ((c ^ (r & 7)) << 4)
