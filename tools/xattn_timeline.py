"""CTA timeline of sta_xattn_fwd (debug build only).  Compiles csrc/sta_xattn_fwd.cu with -DSTA_TIMELINE into
csrc/build/libsta_b200_tl.so (the product library is untouched), runs one geometry and prints, per stamp, the median /
max over CTAs of the SM cycles since the CTA's first instruction.

Stamps: 0 entry | 1 first TMA loads issued | 2 after the prologue __syncthreads | 3 Q_u + K_0 landed | 4 P V(0) issue |
5 P V(last) issue | 6 WG0 S(0) ready | 7 WG1 S(1) ready | 8 WG0 P(0) stored | 9 WG1 P(1) stored | 10 WG0 O_u ready |
11 WG1 O ready | 12 WG0 stores done | 13 WG1 stores done | 14 TMEM freed
"""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from diffusion_spacetime_attn_b200 import build as B  # noqa: E402

NAMES = ["entry", "tma_issued", "post_sync", "q_k0_landed", "pv0_issue", "pv_last_issue", "wg0_s0_ready", "wg1_s1_ready",
         "wg0_p0_stored", "wg1_p1_stored", "wg0_o_ready", "wg1_o_ready", "wg0_stored", "wg1_stored", "tmem_freed"]


def build_debug():
    B.build_native()
    out = B.BUILD / "libsta_b200_tl.so"
    obj = B.BUILD / "sta_xattn_fwd_tl.o"
    src = B.CSRC / "sta_xattn_fwd.cu"
    extra = [f for f in sys.argv[1:] if f.startswith("-D")]
    subprocess.run([B._nvcc(), *B.NVCC_FLAGS, "-DSTA_TIMELINE", *extra, "-I", str(B.INCLUDE), "-c", str(src), "-o", str(obj)],
                   check=True, capture_output=True)
    objs = [str(B.BUILD / (s.stem + ".o")) for s in sorted(B.CSRC.glob("*.cu")) if s.stem != "sta_xattn_fwd"] + [str(obj)]
    subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *objs], check=True)
    return out


if __name__ == "__main__":
    lib = build_debug()
    if "--build-only" in sys.argv:
        print(lib)
        sys.exit(0)
    os.environ["STA_B200_LIB"] = str(lib)
    import torch

    import bench
    from diffusion_spacetime_attn_b200 import native, ops
    from diffusion_spacetime_attn_b200.ldm.modules.attention import build_object_masks

    geoms = [(1, 4096, 8, 40, 2), (1, 1024, 8, 80, 2), (1, 256, 8, 160, 2), (1, 64, 8, 160, 2)]
    h = native.load()
    h.sta_debug_timeline_fwd.argtypes = [C.c_void_p]
    for (Bp, n, heads, d, n_obj) in geoms:
        Cc = heads * d
        g = torch.Generator(device="cuda").manual_seed(0)
        q = torch.randn(2 * Bp, n, Cc, device="cuda", generator=g).half()
        kc = torch.randn(Bp, 2 + n_obj, 77, Cc, device="cuda", generator=g).half()
        vc = torch.randn(Bp, 2 + n_obj, 77, Cc, device="cuda", generator=g).half()
        boxes = [[0.25 + 0.5 * (i % 2), 0.25 + 0.5 * ((i // 2) % 2)] for i in range(n_obj)]
        mask = build_object_masks(boxes, n, "cuda").unsqueeze(0).expand(Bp, -1, -1).contiguous()
        coef = torch.full((Bp, n_obj), 2.5, device="cuda")
        ctas = ((n + 127) // 128) * heads * Bp
        tl = torch.zeros(ctas, 20, dtype=torch.int64, device="cuda")
        for _ in range(3):
            ops.xattn_fwd(q, kc, vc, mask, coef, heads)
        h.sta_debug_timeline_fwd(C.c_void_p(tl.data_ptr()))
        ops.xattn_fwd(q, kc, vc, mask, coef, heads)
        torch.cuda.synchronize()
        h.sta_debug_timeline_fwd(C.c_void_p(0))
        t = tl.cpu()
        rel = (t[:, :15] - t[:, :1]).double()
        gt, gs, sm = t[:, 15], t[:, 16], t[:, 17]
        per_sm = torch.bincount(sm, minlength=148)
        print(f"== xattn_fwd {(Bp, n, heads, d, n_obj)}: {ctas} CTAs on {(per_sm > 0).sum().item()} SMs (max {per_sm.max().item()} per SM); "
              f"globaltimer: start spread {(gs.max() - gs.min()).item()} ns, end spread {(gt.max() - gt.min()).item()} ns, "
              f"first start -> last end {(gt.max() - gs.min()).item()} ns")
        # were two CTAs of one SM resident together?  (second CTA's start before the first one's end)
        overlap = 0
        for m in torch.nonzero(per_sm > 1).flatten().tolist():
            idx = torch.nonzero(sm == m).flatten()
            st, en = gs[idx], gt[idx]
            order = torch.argsort(st)
            if st[order[1]] < en[order[0]]:
                overlap += 1
        print(f"   SMs whose first two CTAs overlapped in time: {overlap} of {(per_sm > 1).sum().item()}")
        c18 = (t[:, 18] - t[:, 0]).double()
        print(f"   18 qk0_issued      median {c18.median().item():8.0f} cyc")
        for i, nm in enumerate(NAMES):
            col = rel[:, i]
            col = col[t[:, i] != 0]
            if col.numel():
                print(f"  {i:2d} {nm:16s} median {col.median().item():8.0f} cyc  max {col.max().item():8.0f} cyc"
                      f"   (~{col.median().item() / 1.9e3:5.2f} us)")
    print("device_error", native.device_error())
