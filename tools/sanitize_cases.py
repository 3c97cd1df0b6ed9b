"""One launch of every tcgen05 / TMA kernel at its smallest supported shape (for compute-sanitizer, see tools/sanitize.sh):
the mbarrier / TMEM / TMA protocols are the same at every size, and the sanitizer slows a launch down 10-100x."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import native, ops  # noqa: E402

which = sys.argv[1:] or ["sattn", "wide", "xattn", "tokens"]
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).half()
if "sattn" in which:
    for (b, n, h, d) in [(1, 200, 2, 40), (1, 136, 1, 80), (1, 130, 1, 160)]:  # ragged last tiles on purpose
        q, k, v, do = rnd(b, n, h * d), rnd(b, n, h * d), rnd(b, n, h * d), rnd(b, n, h * d) * 0.1
        out, lse = ops.sattn_fwd(q, k, v, h)
        ops.sattn_bwd(q, k, v, out, lse, do, h)
        torch.cuda.synchronize()
        print("sattn", (b, n, h, d), "ok", float(out.float().abs().mean()))
if "wide" in which:  # 512-wide head (VAE AttnBlock): two key tiles, ragged, so every ring / pipeline barrier wraps at least once
    for (b, n) in [(1, 200)]:
        q, k, v, do = rnd(b, n, 512), rnd(b, n, 512), rnd(b, n, 512), rnd(b, n, 512) * 0.1
        out, lse = ops.sattn_fwd(q, k, v, 1)
        ops.sattn_bwd(q, k, v, out, lse, do, 1)
        torch.cuda.synchronize()
        print("wide", (b, n, 1, 512), "ok", float(out.float().abs().mean()))
    q, k, v = rnd(10, 1000, 512), rnd(10, 1000, 512), rnd(10, 1000, 512)  # 160 CTAs: the forward's 256-column variant
    out, _ = ops.sattn_fwd(q, k, v, 1)
    torch.cuda.synchronize()
    print("wide fwd256", (10, 1000, 1, 512), "ok", float(out.float().abs().mean()))
if "xattn" in which:
    for (B, n, h, d, n_obj) in [(1, 144, 2, 40, 3), (1, 100, 1, 80, 2), (2, 64, 1, 160, 2)]:
        C = h * d
        q, do = rnd(2 * B, n, C), rnd(2 * B, n, C) * 0.1
        kc, vc = rnd(B, 2 + n_obj, 77, C), rnd(B, 2 + n_obj, 77, C)
        mask = (torch.rand(B, n_obj, n, device="cuda", generator=g) < 0.3).to(torch.uint8)
        coef = torch.full((B, n_obj), 2.5, device="cuda")
        out, lse = ops.xattn_fwd(q, kc, vc, mask, coef, h)
        ops.xattn_bwd(q, kc, vc, mask, coef, lse, do, h, out=out)
        torch.cuda.synchronize()
        print("xattn", (B, n, h, d, n_obj), "ok", float(out.float().abs().mean()))
if "tokens" in which:
    x = rnd(2, 320, 8, 8).contiguous(memory_format=torch.channels_last)
    gam, bet = torch.ones(320, device="cuda"), torch.zeros(320, device="cuda")
    y, stats, xn = ops.groupnorm_fwd(x, gam, bet, 1e-5, True)
    ops.groupnorm_bwd(xn, y, gam, bet, stats, 1e-5, True)
    t = rnd(64, 320)
    s, yy, st = ops.add_layernorm_fwd(t, gam, t, gam, bet, 1e-5)
    ops.add_layernorm_bwd(yy, s, s, st, gam)
    ops.geglu_bwd(rnd(64, 2560), ops.geglu_fwd(rnd(64, 2560)))
    torch.cuda.synchronize()
    print("tokens ok")
print("device_error", native.device_error())
