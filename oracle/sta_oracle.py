"""CPU oracle for the spatial-temporal-attention hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import
this module, and only as the checker.  The product path (diffusion_spacetime_attn_b200/) never imports it and
fails loudly when the CUDA library is missing.

What it is: a functional (state_dict in, tensors out) restatement, in plain torch CPU ops at whatever dtype the
caller passes (fp32 / fp64), of the reference's algorithm for the path named by BASELINE.json `north_star`.
Every function cites the reference file:line it follows; paths are relative to
/root/reference/attention_optimization/stable-diffusion/.

Pinning: the reference's own tests hold NO golden vectors for this path.  The oracle is
therefore pinned against outputs of the UNMODIFIED reference modules imported in the build container by
`oracle/make_golden.py`; the resulting fixtures live in tests/golden/ and `tests/test_oracle_golden.py`
re-checks them on every CPU run.  The sampler arithmetic (`make_schedule`, `plms_update`, `plms_combine`,
`plms_trajectory`) is pinned the same way: `oracle/make_golden_sampler.py` imports the UNMODIFIED reference
ldm/models/diffusion/plms.py behind a `clip` shim, runs its own `make_schedule` + `p_sample_plms` for 5 / 10 / 50 steps on
a linear stand-in UNet, and tests/test_host_cpu.py::test_sampler_pinned_by_execution_of_the_reference_plms compares the
oracle and the product sampler with that fixture (schedule constants, every e_t, final latent).  Gradients: the
reference can only back-propagate under CUDA autocast; tools/ref_on_gpu.py runs it there (baseline/_ref) and
tests/test_ref_gpu_golden.py compares; the oracle's own autograd is pinned by fp64 finite differences.
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

MASK_RADIUS_SQ = 0.04  # r = 0.2 (ldm/modules/attention.py:261)


# ------------------------------------------------------------------------------------------------------
# masks and contexts (ldm/modules/attention.py:240-263)
# ------------------------------------------------------------------------------------------------------
def disc_mask(center_xy: Sequence[float], dim: int) -> Tensor:
    """Boolean [dim, dim] mask of latent pixels inside the disc of radius 0.2 around (x, y).

    attention.py:250-261: axis = arange(dim, fp32)/dim; dist[r, c] = (axis[c]-x)^2 + (axis[r]-y)^2 < 0.04.
    Row index is the y direction (`dist2.unsqueeze(1)`), column index the x direction.
    """
    axis = torch.arange(dim, dtype=torch.float32) / dim
    dx = (axis - center_xy[0]) ** 2
    dy = (axis - center_xy[1]) ** 2
    return (dx.unsqueeze(0) + dy.unsqueeze(1)) < MASK_RADIUS_SQ


def flat_masks(bboxes: Sequence[Sequence[float]], n_tokens: int) -> Tensor:
    """uint8 [n_obj, n_tokens]; tokens are row-major (h w) as in 'b c h w -> b (h w) c' (attention.py:341)."""
    dim = int(math.isqrt(n_tokens))
    assert dim * dim == n_tokens, "the reference assumes a square latent (attention.py:243)"
    if len(bboxes) == 0:
        return torch.zeros(0, n_tokens, dtype=torch.uint8)
    return torch.stack([disc_mask(b, dim).reshape(-1) for b in bboxes]).to(torch.uint8)


# ------------------------------------------------------------------------------------------------------
# attention primitives (ldm/modules/attention.py:157-215)
# ------------------------------------------------------------------------------------------------------
def _split_heads(t: Tensor, heads: int) -> Tensor:  # 'b n (h d) -> b h n d'  (attention.py:183)
    b, n, c = t.shape
    return t.reshape(b, n, heads, c // heads).permute(0, 2, 1, 3)


def _merge_heads(t: Tensor) -> Tensor:  # 'b h n d -> b n (h d)'  (attention.py:197)
    b, h, n, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b, n, h * d)


def attention_core(q: Tensor, k: Tensor, v: Tensor, heads: int, return_lse: bool = False):
    """softmax(q k^T * d^-1/2) v per head, merged back (attention.py:183-197).  q [b,n,C], k/v [b,m,C]."""
    d = q.shape[-1] // heads
    qh, kh, vh = (_split_heads(t, heads) for t in (q, k, v))
    sim = torch.einsum("bhid,bhjd->bhij", qh, kh) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = _merge_heads(torch.einsum("bhij,bhjd->bhid", attn, vh))
    if return_lse:
        return out, torch.logsumexp(sim, dim=-1)  # [b, h, n]
    return out


def cross_attention(x: Tensor, context: Optional[Tensor], p: StateDict, prefix: str, heads: int) -> Tensor:
    """CrossAttention.forward (attention.py:175-215) without the dead mask/self_attention_region branches."""
    ctx = x if context is None else context
    q = F.linear(x, p[prefix + "to_q.weight"])
    k = F.linear(ctx, p[prefix + "to_k.weight"])
    v = F.linear(ctx, p[prefix + "to_v.weight"])
    o = attention_core(q, k, v, heads)
    return F.linear(o, p[prefix + "to_out.0.weight"], p[prefix + "to_out.0.bias"])


def feed_forward(x: Tensor, p: StateDict, prefix: str) -> Tensor:
    """GEGLU feed-forward (attention.py:42-69): Linear(C, 8C) -> a * gelu(gate) -> Linear(4C, C)."""
    h = F.linear(x, p[prefix + "net.0.proj.weight"], p[prefix + "net.0.proj.bias"])
    a, gate = h.chunk(2, dim=-1)
    return F.linear(a * F.gelu(gate), p[prefix + "net.2.weight"], p[prefix + "net.2.bias"])


def _layer_norm(x: Tensor, p: StateDict, prefix: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), p[prefix + "weight"], p[prefix + "bias"], 1e-5)


# ------------------------------------------------------------------------------------------------------
# the hot path: BasicTransformerBlock._forward (ldm/modules/attention.py:268-300)
# ------------------------------------------------------------------------------------------------------
def transformer_block(
    x: Tensor,
    context: Tensor,
    coef: Tensor,
    local_contexts: Sequence[Tensor],
    masks: Tensor,
    p: StateDict,
    prefix: str,
    heads: int,
) -> Tensor:
    """Out-of-place restatement of the reference block for a batch of B prompts.

    x [2B, N, C] (rows [0,B) unconditional, [B,2B) conditional — c_in = cat([uc, c]), plms.py:306),
    context [2B, 77, 768], coef [n_obj] or [B, n_obj], local_contexts: n_obj tensors [2B, 77, 768] — what the
    reference calls curr_cs[i] = cat(self.uncond, c_i) (attention.py:246-248), masks uint8/bool [n_obj, N] or
    [B, n_obj, N].

    Reference (B = 1): gs_i = attn2(norm2(h), curr_cs[i]); g = attn2(norm2(h), context);
    o[0] = g[0]; o[1] = g[1] + sum_i mask_i * ((coef_i*gs_i)[1] - (coef_i*g)[0])   (attention.py:278-294).
    It subtracts the UNCONDITIONAL row of the global output, which is what is restated here.
    """
    B = x.shape[0] // 2
    n_obj = len(local_contexts)
    h = cross_attention(_layer_norm(x, p, prefix + "norm1."), None, p, prefix + "attn1.", heads) + x  # :274
    hn = _layer_norm(h, p, prefix + "norm2.")
    gs = [cross_attention(hn, lc, p, prefix + "attn2.", heads) for lc in local_contexts]  # :278-279
    g = cross_attention(hn, context, p, prefix + "attn2.", heads)  # :281
    o_u = g[:B]
    o_c = g[B:]
    if n_obj:
        coef_b = coef.reshape(-1, n_obj).expand(B, n_obj).to(x.dtype)
        m = masks.reshape(-1, n_obj, x.shape[1]).expand(B, n_obj, x.shape[1]).to(x.dtype)
        for i in range(n_obj):  # :284-294
            c_i = coef_b[:, i].reshape(B, 1, 1)
            diff = c_i * gs[i][B:] - c_i * g[:B]
            o_c = o_c + m[:, i].unsqueeze(-1) * diff
    z = torch.cat([o_u, o_c], dim=0) + h  # :297
    return feed_forward(_layer_norm(z, p, prefix + "norm3."), p, prefix + "ff.") + z  # :299


def dual_cross_attention_core(
    q: Tensor, k_ctx: Tensor, v_ctx: Tensor, masks: Tensor, coef: Tensor, heads: int, return_lse: bool = False
):
    """The quantity the fused CUDA kernel (sta_xattn_fwd) produces: the blend moved BEFORE to_out.

    q [2B, N, C]; k_ctx / v_ctx [B, 2+n_obj, L, C] (slot 0 uncond ctx, 1 global, 2+i local i);
    masks [B, n_obj, N]; coef [B, n_obj].  Returns [2B, N, C]:
      out[b]   = A_u,   out[B+b] = A_g + sum_i m_i c_i (A_i - A_u)
    Because mask and coef are per-pixel scalars and to_out is affine, to_out(out) + residual equals the
    reference block's blend (checked to 1e-15 in tests/test_oracle_golden.py).
    """
    B = q.shape[0] // 2
    n_obj = k_ctx.shape[1] - 2
    q_u, q_c = q[:B], q[B:]
    lses = []

    def att(qq, slot):
        r = attention_core(qq, k_ctx[:, slot], v_ctx[:, slot], heads, return_lse=True)
        lses.append(r[1])
        return r[0]

    a_u = att(q_u, 0)
    a_g = att(q_c, 1)
    out_c = a_g
    for i in range(n_obj):
        a_i = att(q_c, 2 + i)
        s = (masks[:, i].to(q.dtype) * coef[:, i].to(q.dtype).unsqueeze(-1)).unsqueeze(-1)  # [B, N, 1]
        out_c = out_c + s * (a_i - a_u)
    out = torch.cat([a_u, out_c], dim=0)
    if return_lse:
        return out, torch.stack(lses, dim=2)  # [B, h, 2+n_obj, N]
    return out


# ------------------------------------------------------------------------------------------------------
# SpatialTransformer / ResBlock / UNet (attention.py:335-345, openaimodel.py:163-275, 443-742)
# ------------------------------------------------------------------------------------------------------
def _group_norm32(x: Tensor, p: StateDict, prefix: str, eps: float) -> Tensor:
    # GroupNorm32 computes in fp32 and casts back (util.py:214-216); SpatialTransformer uses eps 1e-6
    # (attention.py:79), the UNet's normalization() uses the nn.GroupNorm default 1e-5.
    return F.group_norm(x.float(), 32, p[prefix + "weight"].float(), p[prefix + "bias"].float(), eps).to(x.dtype)


def spatial_transformer(x, context, coef, local_contexts, bboxes, p, prefix, heads) -> Tensor:
    """SpatialTransformer.forward (attention.py:335-345), depth 1."""
    b, c, hh, ww = x.shape
    x_in = x
    y = _group_norm32(x, p, prefix + "norm.", 1e-6)
    y = F.conv2d(y, p[prefix + "proj_in.weight"], p[prefix + "proj_in.bias"])
    y = y.permute(0, 2, 3, 1).reshape(b, hh * ww, -1)
    masks = flat_masks(bboxes, hh * ww)
    y = transformer_block(y, context, coef, local_contexts, masks, p, prefix + "transformer_blocks.0.", heads)
    y = y.reshape(b, hh, ww, -1).permute(0, 3, 1, 2)
    y = F.conv2d(y, p[prefix + "proj_out.weight"], p[prefix + "proj_out.bias"])
    return y + x_in


def res_block(x: Tensor, emb: Tensor, p: StateDict, prefix: str) -> Tensor:
    """ResBlock._forward without up/down and scale-shift (openaimodel.py:252-275, SD-v1 config)."""
    h = _group_norm32(x, p, prefix + "in_layers.0.", 1e-5)
    h = F.conv2d(F.silu(h), p[prefix + "in_layers.2.weight"], p[prefix + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), p[prefix + "emb_layers.1.weight"], p[prefix + "emb_layers.1.bias"]).to(h.dtype)
    h = h + e[:, :, None, None]
    h = _group_norm32(h, p, prefix + "out_layers.0.", 1e-5)
    h = F.conv2d(F.silu(h), p[prefix + "out_layers.3.weight"], p[prefix + "out_layers.3.bias"], padding=1)
    if prefix + "skip_connection.weight" in p:
        x = F.conv2d(x, p[prefix + "skip_connection.weight"], p[prefix + "skip_connection.bias"])
    return x + h


def timestep_embedding(timesteps: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """Sinusoidal embedding, cos first then sin (util.py:151-171)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


@dataclass
class UNetConfig:
    """SD-v1 UNet hyper-parameters (configs/stable-diffusion/v1-inference.yaml:29-44)."""

    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    attention_resolutions: Tuple[int, ...] = (4, 2, 1)
    num_res_blocks: int = 2
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_heads: int = 8
    context_dim: int = 768

    def layout(self) -> Tuple[List[List[Tuple[str, int, int]]], List[Tuple[str, int, int]], List[List[Tuple[str, int, int]]]]:
        """Stage lists mirroring UNetModel.__init__ (openaimodel.py:521-673).

        Each stage is a list of (kind, in_ch, out_ch), kind in {conv, res, attn, down, up}.
        """
        mc = self.model_channels
        inp: List[List[Tuple[str, int, int]]] = [[("conv", self.in_channels, mc)]]
        chans = [mc]
        ch, ds = mc, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(self.num_res_blocks):
                st = [("res", ch, mult * mc)]
                ch = mult * mc
                if ds in self.attention_resolutions:
                    st.append(("attn", ch, ch))
                inp.append(st)
                chans.append(ch)
            if level != len(self.channel_mult) - 1:
                inp.append([("down", ch, ch)])
                chans.append(ch)
                ds *= 2
        mid = [("res", ch, ch), ("attn", ch, ch), ("res", ch, ch)]
        out: List[List[Tuple[str, int, int]]] = []
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(self.num_res_blocks + 1):
                ich = chans.pop()
                st = [("res", ch + ich, mc * mult)]
                ch = mc * mult
                if ds in self.attention_resolutions:
                    st.append(("attn", ch, ch))
                if level and i == self.num_res_blocks:
                    st.append(("up", ch, ch))
                    ds //= 2
                out.append(st)
        return inp, mid, out


def unet_param_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape of the reference UNetModel for `cfg` (checked against the real module's
    state_dict in oracle/make_golden.py; fixture tests/golden/unet_tiny_shapes.json)."""
    s: Dict[str, Tuple[int, ...]] = {}
    mc, ted, cd = cfg.model_channels, cfg.model_channels * 4, cfg.context_dim

    def lin(name, o, i, bias=True):
        s[name + ".weight"] = (o, i)
        if bias:
            s[name + ".bias"] = (o,)

    def conv(name, o, i, k):
        s[name + ".weight"] = (o, i, k, k)
        s[name + ".bias"] = (o,)

    def norm(name, c):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)

    def res(pre, i, o):
        norm(pre + "in_layers.0", i)
        conv(pre + "in_layers.2", o, i, 3)
        lin(pre + "emb_layers.1", o, ted)
        norm(pre + "out_layers.0", o)
        conv(pre + "out_layers.3", o, o, 3)
        if i != o:
            conv(pre + "skip_connection", o, i, 1)

    def attn(pre, c):
        norm(pre + "norm", c)
        conv(pre + "proj_in", c, c, 1)
        t = pre + "transformer_blocks.0."
        for a, ctx in (("attn1", c), ("attn2", cd)):
            lin(t + a + ".to_q", c, c, bias=False)
            lin(t + a + ".to_k", c, ctx, bias=False)
            lin(t + a + ".to_v", c, ctx, bias=False)
            lin(t + a + ".to_out.0", c, c)
        lin(t + "ff.net.0.proj", 8 * c, c)
        lin(t + "ff.net.2", c, 4 * c)
        for k in ("norm1", "norm2", "norm3"):
            norm(t + k, c)
        conv(pre + "proj_out", c, c, 1)

    lin("time_embed.0", ted, mc)
    lin("time_embed.2", ted, ted)
    inp, mid, out = cfg.layout()

    def stage(pre, st):
        for j, (kind, i, o) in enumerate(st):
            q = f"{pre}{j}."
            if kind == "conv":
                conv(q[:-1], o, i, 3)
            elif kind == "res":
                res(q, i, o)
            elif kind == "attn":
                attn(q, o)
            elif kind == "down":
                conv(q + "op", o, i, 3)
            elif kind == "up":
                conv(q + "conv", o, i, 3)

    for n, st in enumerate(inp):
        stage(f"input_blocks.{n}.", st)
    stage("middle_block.", mid)
    for n, st in enumerate(out):
        stage(f"output_blocks.{n}.", st)
    norm("out.0", mc)
    conv("out.2", cfg.out_channels, mc, 3)
    return s


def seeded_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int, dtype=torch.float32) -> StateDict:
    """Deterministic random weights (no checkpoint is available offline).

    Every tensor gets its own generator seeded by crc32(key) ^ seed, so the values do not depend on iteration
    order.  The reference's zero_module()'d layers (attention.py:329, openaimodel.py:229-231, 685) are
    randomised like any other — with zeros there the UNet output is identically 0 and parity would be vacuous.
    Scales: matrices/convs ~ N(0, 1/fan_in), biases ~ N(0, 0.02^2), norm weights ~ 1 + N(0, 0.1^2).
    """
    sd: StateDict = {}
    for key, shape in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        is_norm = ".norm" in key or key.startswith("out.0") or "in_layers.0" in key or "out_layers.0" in key
        if key.endswith(".bias"):
            t = t * 0.02
        elif is_norm:
            t = 1.0 + 0.1 * t
        else:
            fan_in = int(np.prod(shape[1:]))
            t = t / math.sqrt(fan_in)
        sd[key] = t.to(dtype)
    return sd


def unet_forward(
    x: Tensor,
    timesteps: Tensor,
    context: Tensor,
    coef: Tensor,
    bboxes: Sequence[Sequence[float]],
    local_contexts: Sequence[Tensor],
    p: StateDict,
    cfg: UNetConfig,
) -> Tensor:
    """UNetModel.forward (openaimodel.py:710-742) for the crossattn conditioning key.

    x [2B,4,H,W], timesteps [2B], context [2B,77,768], local_contexts: n_obj x [2B,77,768].
    """
    emb = timestep_embedding(timesteps, cfg.model_channels).to(x.dtype)
    emb = F.linear(emb, p["time_embed.0.weight"], p["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), p["time_embed.2.weight"], p["time_embed.2.bias"])
    inp, mid, out = cfg.layout()

    def run(pre, st, h):
        for j, (kind, _i, _o) in enumerate(st):
            q = f"{pre}{j}."
            if kind == "conv":
                h = F.conv2d(h, p[q + "weight"], p[q + "bias"], padding=1)
            elif kind == "res":
                h = res_block(h, emb, p, q)
            elif kind == "attn":
                h = spatial_transformer(h, context, coef, local_contexts, bboxes, p, q, cfg.num_heads)
            elif kind == "down":
                h = F.conv2d(h, p[q + "op.weight"], p[q + "op.bias"], stride=2, padding=1)
            elif kind == "up":
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, p[q + "conv.weight"], p[q + "conv.bias"], padding=1)
        return h

    hs = []
    h = x
    for n, st in enumerate(inp):
        h = run(f"input_blocks.{n}.", st, h)
        hs.append(h)
    h = run("middle_block.", mid, h)
    for n, st in enumerate(out):
        h = run(f"output_blocks.{n}.", st, torch.cat([h, hs.pop()], dim=1))
    h = F.silu(_group_norm32(h, p, "out.0.", 1e-5))
    return F.conv2d(h, p["out.2.weight"], p["out.2.bias"], padding=1)


# ------------------------------------------------------------------------------------------------------
# noise schedule and PLMS sampler (util.py:21-74, ddpm.py:116-130, plms.py:81-112, 296-358)
# ------------------------------------------------------------------------------------------------------
@dataclass
class Schedule:
    timesteps: np.ndarray  # ddim_timesteps, ascending, e.g. [1, 21, ..., 981] for S = 50
    alphas: np.ndarray
    alphas_prev: np.ndarray
    sqrt_one_minus_alphas: np.ndarray
    sigmas: np.ndarray


def make_schedule(num_steps: int, num_ddpm: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.012) -> Schedule:
    """betas linear in sqrt space (util.py:21-25 with v1-inference.yaml:5-6), alphas_cumprod in fp64 cast to
    fp32 (ddpm.py:124-133), uniform DDIM sub-sampling +1 (util.py:46-60), eta = 0 (plms.py:82-83)."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_ddpm, dtype=torch.float64).numpy() ** 2
    acp = np.cumprod(1.0 - betas, axis=0).astype(np.float32)  # register_buffer(..., float32)
    c = num_ddpm // num_steps
    ts = np.asarray(list(range(0, num_ddpm, c))) + 1
    alphas = acp[ts]
    alphas_prev = np.asarray([acp[0]] + acp[ts[:-1]].tolist())
    sigmas = 0.0 * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return Schedule(ts, alphas, alphas_prev, np.sqrt(1.0 - alphas), sigmas)


def plms_update(x: Tensor, e_t: Tensor, sch: Schedule, index: int) -> Tuple[Tensor, Tensor]:
    """get_x_prev_and_pred_x0 (plms.py:321-338) with eta = 0 (no noise term)."""
    a_t = torch.tensor(float(sch.alphas[index]), dtype=x.dtype)
    a_prev = torch.tensor(float(sch.alphas_prev[index]), dtype=x.dtype)
    sigma_t = torch.tensor(float(sch.sigmas[index]), dtype=x.dtype)
    s1m = torch.tensor(float(sch.sqrt_one_minus_alphas[index]), dtype=x.dtype)
    pred_x0 = (x - s1m * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
    return a_prev.sqrt() * pred_x0 + dir_xt, pred_x0


def plms_combine(e_t: Tensor, old_eps: List[Tensor]) -> Tensor:
    """Adams-Bashforth combination for len(old_eps) >= 1 (plms.py:346-354)."""
    if len(old_eps) == 1:
        return (3 * e_t - old_eps[-1]) / 2
    if len(old_eps) == 2:
        return (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
    return (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24


def plms_trajectory(
    eps_model: Callable[[Tensor, int, int], Tensor],
    x_T: Tensor,
    num_steps: int,
    sch: Optional[Schedule] = None,
) -> Tensor:
    """The 50-step loop of plms_sampling (plms.py:224-247) around p_sample_plms (plms.py:296-358).

    eps_model(x, t, step_i) returns the guided eps for latent x at DDPM timestep t; step_i is the loop index
    i (the column of weighting_parameter the reference passes as coef, plms.py:243).  At i = 0 the model is
    evaluated a second time at (x_prev, t_next) with the SAME step index (plms.py:341-345).
    """
    sch = sch or make_schedule(num_steps)
    time_range = np.flip(sch.timesteps)
    total = len(time_range)
    old_eps: List[Tensor] = []
    img = x_T
    for i, step in enumerate(time_range):
        index = total - i - 1
        t_next = int(time_range[min(i + 1, total - 1)])
        e_t = eps_model(img, int(step), i)
        if len(old_eps) == 0:
            x_prev, _ = plms_update(img, e_t, sch, index)
            e_next = eps_model(x_prev, t_next, i)
            e_prime = (e_t + e_next) / 2
        else:
            e_prime = plms_combine(e_t, old_eps)
        img, _ = plms_update(img, e_prime, sch, index)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
    return img


def guided_eps(
    x: Tensor, t: int, coef: Tensor, uc: Tensor, c: Tensor, local_cs: Sequence[Tensor], uncond_block: Tensor,
    bboxes, p: StateDict, cfg: UNetConfig, scale: float = 7.5,
) -> Tensor:
    """get_model_output (plms.py:299-308): CFG batch [uncond, cond] through apply_model_extra
    (ddpm.py:891-905 -> 1420-1428), e = e_u + s (e_c - e_u).  B = 1 as in the reference.

    local_cs[i] is c_i [1,77,768]; the block pairs it with its own stored uncond embedding
    (curr_cs[i] = cat(self.uncond, c_i), attention.py:246-247) — `uncond_block` is that tensor.
    """
    x_in = torch.cat([x, x])
    t_in = torch.full((2,), t, dtype=torch.long)
    c_in = torch.cat([uc, c])
    locals_ = [torch.cat([uncond_block.to(x.dtype), lc]) for lc in local_cs]
    e_u, e_c = unet_forward(x_in, t_in, c_in, coef, bboxes, locals_, p, cfg).chunk(2)
    return e_u + scale * (e_c - e_u)


def alpha_init(n_obj: int, num_steps: int) -> Tensor:
    """weighting_parameter = 5 / n_obj everywhere (plms.py:204-210; the reference hard-codes 50 columns)."""
    return torch.full((n_obj, num_steps), 5.0 / max(n_obj, 1), dtype=torch.float32)


def object_crop_box(center: Sequence[float], size: int = 512) -> Tuple[int, int, int, int]:
    """Pixel crop (y1, y2, x1, x2) of the per-object CLIP loss: centre +- 0.2 clamped to [0,1] (plms.py:256-270)."""
    x1, x2 = max(center[0] - 0.2, 0), min(center[0] + 0.2, 1)
    y1, y2 = max(center[1] - 0.2, 0), min(center[1] + 0.2, 1)
    return int(size * y1), int(size * y2), int(size * x1), int(size * x2)
