"""Time the 512-wide flash attention (sta_sattn_wide.cu) against the materialised bmm / softmax / bmm formulation of the
KL-VAE mid-block AttnBlock (model.py:176-191) on the same inputs: forward and forward+backward, CUDA events, 10 launches
after 3 warm-ups.  Usage: python tools/bench_wide.py [n batch]..."""
import sys

import torch

sys.path.insert(0, ".")
from diffusion_spacetime_attn_b200 import native, ops  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    cases = [(4096, 1), (9216, 2)]
    if len(sys.argv) > 2:
        cases = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
    c = 512
    for n, b in cases:
        qkv = torch.randn(b, n, 3 * c, device="cuda").half()
        d_out = (torch.randn(b, n, c, device="cuda") * 0.1).half()
        q, k, v = qkv.chunk(3, dim=-1)

        def flash_fwd():
            return ops.sattn_fwd(q, k, v, 1)

        out, lse = flash_fwd()

        def flash_bwd():
            return ops.sattn_bwd(q, k, v, out, lse, d_out, 1)

        def mat_fwd():
            w = torch.bmm(q, k.transpose(1, 2))
            w = torch.softmax(w.float() * (c ** -0.5), dim=2).half()
            return torch.bmm(w, v)

        qkv_g = qkv.clone().requires_grad_(True)

        def mat_fwd_bwd():
            qg, kg, vg = qkv_g.chunk(3, dim=-1)
            w = torch.bmm(qg, kg.transpose(1, 2))
            w = torch.softmax(w.float() * (c ** -0.5), dim=2).half()
            o = torch.bmm(w, vg)
            qkv_g.grad = None
            o.backward(d_out)

        t_ff, t_fb = timed(flash_fwd), timed(flash_bwd)
        t_mf, t_mfb = timed(mat_fwd), timed(mat_fwd_bwd)
        flops = 4.0 * n * n * c * b
        print(f"n={n} b={b}: flash fwd {t_ff:.1f} us ({flops / t_ff * 1e-6:.0f} TFLOP/s alg), flash bwd {t_fb:.1f} us "
              f"({2.5 * flops / t_fb * 1e-6:.0f} TFLOP/s alg) | materialised fwd {t_mf:.1f} us, fwd+bwd {t_mfb:.1f} us "
              f"(bwd ~{t_mfb - t_mf:.1f})", flush=True)
        assert native.device_error() == 0


if __name__ == "__main__":
    main()
