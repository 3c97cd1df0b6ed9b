"""Public API of the package: one object that owns the frozen networks on one GPU and turns (prompt, layout) work
items into images through the drop-in `ldm` modules — what scripts/txt2img-*.py, bench.py and smoke() call.

Mirrors the per-prompt body of the reference entrypoints (scripts/txt2img-gpt.py:303-341): seed, layout -> bboxes,
text stage -> (uc, c, c_i), `sampler.sample(...)`.  Differences: embeddings travel in memory, not through
c{i}_*.pt files; several prompts can share one call (B > 1); weights come from a checkpoint when one is given
(`state_dict` keys of the reference's LatentDiffusion) and are seeded random tensors otherwise (no checkpoint exists
offline: SURVEY.md §0).

Multi-GPU: one process per GPU; `shard_prompts` gives rank r the prompt indices {i : i mod W == r} (the reference
shards by hand-edited `start` ranges, txt2img-gpt.py:303-305).  The only collective is a broadcast of the frozen
weights from rank 0 at start-up (`broadcast_weights`).
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .ldm.models.diffusion.ddim import DDIMSampler
from .ldm.models.diffusion.ddpm import V1_UNET, V1_VAE, LatentDiffusion
from .ldm.models.diffusion.plms import PLMSSampler
from .ldm.modules.encoders.clip_loss import DCLIPLoss
from .ldm.modules.encoders.modules import SyntheticTextEmbedder


@dataclass
class WorkItem:
    """One prompt with its layout (the reference's `result = inference_sentence(prompt)` dict, txt2img-gpt.py:307)."""

    prompt: str
    prompt_idx: int
    object_names: List[str]
    bboxes: List[List[float]]  # [x_center, y_center] in [0, 1] per object
    seed: int = 1


def shard_prompts(n_items: int, rank: int, world_size: int) -> List[int]:
    """Prompt indices of `rank`: round-robin, so every rank keeps the GLOBAL prompt index in its output file names."""
    return list(range(rank, n_items, world_size))


def synthetic_layout(names: Sequence[str], prompt: str) -> List[List[float]]:
    """Deterministic stand-in for the layout predictor (LayoutTransformer needs fairseq/spaCy + a checkpoint, all
    absent offline): objects on a grid, jittered by a hash of the prompt, centres kept inside [0.2, 0.8]."""
    n = len(names)
    cols = max(1, math.ceil(math.sqrt(n)))
    rows = math.ceil(n / cols)
    out = []
    for i, nm in enumerate(names):
        r, c = divmod(i, cols)
        jx = (zlib.crc32((prompt + nm + "x").encode()) % 1000) / 1000.0 - 0.5
        jy = (zlib.crc32((prompt + nm + "y").encode()) % 1000) / 1000.0 - 0.5
        x = 0.2 + 0.6 * ((c + 0.5) / cols) + 0.1 * jx / cols
        y = 0.2 + 0.6 * ((r + 0.5) / rows) + 0.1 * jy / rows
        out.append([round(min(max(x, 0.0), 1.0), 3), round(min(max(y, 0.0), 1.0), 3)])
    return out


def randomize_zero_modules(model: torch.nn.Module, std: float = 0.02, seed: int = 0) -> None:
    """The reference zero-initialises proj_out / ResBlock out-convs / the final conv (attention.py:329,
    openaimodel.py:229-231,685).  With random (not trained) weights that makes the UNet output identically zero, so
    benchmarks and parity runs re-randomise every all-zero parameter."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            if p.numel() and float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * std)


def broadcast_weights(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 256 << 20) -> int:
    """NCCL broadcast of every parameter and buffer from `src` (start-up only; nothing is exchanged per step).
    Tensors are packed into flat buckets so that ~4 GB of weights need tens of collectives, not thousands."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    tensors = [t for t in list(module.parameters()) + list(module.buffers()) if t.is_floating_point()]
    sent, bucket, size = 0, [], 0

    def flush():
        nonlocal bucket, size, sent
        if not bucket:
            return
        flat = torch.cat([t.detach().reshape(-1).float() for t in bucket])
        dist.broadcast(flat, src=src)
        off = 0
        with torch.no_grad():
            for t in bucket:
                t.copy_(flat[off:off + t.numel()].reshape(t.shape))
                off += t.numel()
        sent += flat.numel() * 4
        bucket, size = [], 0

    for t in tensors:
        bucket.append(t)
        size += t.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return sent


class SpaceTimeAttnPipeline:
    def __init__(self, device="cuda", ckpt: Optional[str] = None, seed: int = 0, steps: int = 50, scale: float = 7.5,
                 latent_size: int = 64, sampler: str = "plms", use_checkpoint: bool = True,
                 checkpoint_min_tokens: int = 0, num_epochs: int = 3, save_images: bool = False,
                 out_dir: str = "result_outputs", with_vae: bool = True, unet_config: Optional[dict] = None,
                 cuda_graphs: bool = True, half_weights: bool = True, allow_synthetic_conditioning: Optional[bool] = None):
        """`ckpt` = a reference checkpoint (`state_dict` keys of LatentDiffusion).  With a checkpoint the text stage and the
        CLIP loss must be real too: `$STA_CLIP_L_PATH` (HF CLIP-L/14 text model + tokenizer directory; the checkpoint's own
        `cond_stage_model.*` tensors are loaded over it) and `$STA_CLIP_B32_PATH` (OpenAI ViT-B/32 state dict) +
        `$STA_CLIP_TOKENIZER_PATH` (CLIP BPE tokenizer directory).  If one of them is missing the constructor RAISES unless
        `allow_synthetic_conditioning=True` — trained UNet weights conditioned on hashed noise and optimised against a
        random-weight loss give meaningless images.  Without a checkpoint everything is synthetic by construction
        (`self.data == "synthetic"`)."""
        from . import native

        native.load()  # fail loudly before building 4 GB of networks if the CUDA library is missing
        # fixed shapes for 150+ evaluations per image: let cuDNN pick its best kernels (off under ncu: the autotuner's
        # trial launches would all be serialised by the profiler)
        import os

        torch.backends.cudnn.benchmark = os.environ.get("STA_CUDNN_BENCHMARK", "1") == "1"
        if "STA_CUDNN_BENCHMARK_LIMIT" in os.environ:  # candidates the autotuner times per convolution (torch default 10, 0 = all)
            torch.backends.cudnn.benchmark_limit = int(os.environ["STA_CUDNN_BENCHMARK_LIMIT"])
        self.device = torch.device(device)
        self.steps, self.scale, self.latent_size = steps, scale, latent_size
        torch.manual_seed(seed)
        cfg = dict(unet_config or V1_UNET)
        cfg["use_checkpoint"] = use_checkpoint
        self.model = LatentDiffusion(unet_config={"params": cfg}, first_stage_config={"params": V1_VAE},
                                     build_first_stage=with_vae)
        if ckpt:
            sd = torch.load(ckpt, map_location="cpu")
            sd = sd.get("state_dict", sd)
            missing, unexpected = self.model.load_state_dict(sd, strict=False)
            self.weights = f"checkpoint:{ckpt} (missing {len(missing)}, unexpected {len(unexpected)})"
        else:
            randomize_zero_modules(self.model, seed=seed)
            self.weights = f"seeded-random(seed={seed})"
        self.model.to(self.device).eval().requires_grad_(False)
        unet = self.model.model.diffusion_model
        if half_weights:
            # fp16 conv/linear weights, fp32 norms: numerically what torch.autocast computes from fp32 masters
            # (scripts/txt2img-gpt.py:310), without re-casting 3.4 GB of weights on every evaluation
            unet.half()
            for m in unet.modules():
                if isinstance(m, (torch.nn.GroupNorm, torch.nn.LayerNorm)):
                    m.float()
            # NHWC activations/weights: cuDNN's fp16 tensor-core convolutions are NHWC natively (removes the
            # nchwToNhwc transposes, ~20 % of an evaluation) and 'b c h w -> b (h w) c' becomes a free view
            unet.to(memory_format=torch.channels_last)
            vae = self.model.first_stage_model
            if vae is not None:  # same treatment for the KL-VAE decoder (fused NHWC GroupNorm / bias / residual kernels)
                vae.half()
                for m in vae.modules():
                    if isinstance(m, torch.nn.GroupNorm):
                        m.float()
                vae.to(memory_format=torch.channels_last)
        self.cuda_graphs = cuda_graphs
        if cuda_graphs:
            from .graphed import GraphedModelRunner

            unet.set_checkpointing(False)  # evaluation-level recompute lives in graphed.py
            self.model.graph_runner = GraphedModelRunner(unet)
        else:
            unet.set_checkpointing(use_checkpoint, checkpoint_min_tokens)
        self.data = "synthetic" if not ckpt else "checkpoint"
        self.text = None
        if ckpt:
            self.text = self._real_text_stage(sd, allow_synthetic_conditioning)
        if self.text is None:
            self.text = SyntheticTextEmbedder(device="cpu")
        self.model.cond_stage_model = self.text
        self.clip_loss = torch.nn.Identity()
        if with_vae:
            self.clip_loss = self._real_clip_loss(allow_synthetic_conditioning) if ckpt else None
            if self.clip_loss is None:
                self.clip_loss = DCLIPLoss(device=self.device, seed=seed + 1)
        if half_weights and with_vae:
            # fp16 CLIP weights (fp32 LayerNorms): what autocast computes from fp32 masters, minus ~460 cast launches
            # per loss evaluation.  The text features are cached per prompt in fp32.
            clip = self.clip_loss.model
            clip.half()
            for m in clip.modules():
                if isinstance(m, torch.nn.LayerNorm):
                    m.float()
        if cuda_graphs and with_vae:
            self.clip_loss.graph_encode = True
        cls = DDIMSampler if sampler == "ddim" else PLMSSampler
        self.sampler = cls(self.model, clip_loss_model=self.clip_loss, num_epochs=num_epochs, save_images=save_images,
                           out_dir=out_dir)

    # -- real conditioning when a checkpoint is given (never a silent fallback) -------------------------------
    def _synthetic_or_raise(self, what: str, hint: str, allow: Optional[bool]):
        msg = (f"a checkpoint was given but {what} is not available ({hint}); the images would be conditioned on / "
               "optimised against synthetic stand-ins")
        if not allow:
            raise RuntimeError(msg + " — pass allow_synthetic_conditioning=True (--allow_synthetic_conditioning) to accept that")
        import warnings

        warnings.warn(msg, stacklevel=3)
        self.data = "checkpoint + SYNTHETIC conditioning"
        return None

    def _real_text_stage(self, sd, allow):
        import os

        from .ldm.modules.encoders.modules import FrozenCLIPEmbedder

        path = os.environ.get("STA_CLIP_L_PATH", "")
        if not os.path.isdir(path):
            return self._synthetic_or_raise("the CLIP-L/14 text stage", "set STA_CLIP_L_PATH to a local HF model directory", allow)
        text = FrozenCLIPEmbedder(path, device="cpu")
        own = {k[len("cond_stage_model."):]: v for k, v in sd.items() if k.startswith("cond_stage_model.")}
        if own:  # the checkpoint's own text tower wins over the directory's weights
            missing, unexpected = text.load_state_dict(own, strict=False)
            if [k for k in missing if "position_ids" not in k]:
                raise RuntimeError(f"cond_stage_model.* of the checkpoint does not fit CLIPTextModel: missing {missing[:5]}")
        return text

    def _real_clip_loss(self, allow):
        import os

        b32, tok = os.environ.get("STA_CLIP_B32_PATH", ""), os.environ.get("STA_CLIP_TOKENIZER_PATH", os.environ.get("STA_CLIP_L_PATH", ""))
        if not os.path.isfile(b32) or not os.path.isdir(tok):
            return self._synthetic_or_raise("the CLIP ViT-B/32 loss", "set STA_CLIP_B32_PATH (OpenAI ViT-B-32 state dict) and "
                                            "STA_CLIP_TOKENIZER_PATH (CLIP BPE tokenizer directory)", allow)
        from transformers import CLIPTokenizer

        tokenizer = CLIPTokenizer.from_pretrained(tok)

        def tokenize(texts):  # clip.tokenize: <sot> ids <eot>, context length 77 (padding beyond <eot> is never attended)
            return tokenizer(list(texts), truncation=True, max_length=77, padding="max_length", return_tensors="pt")["input_ids"]

        loss = DCLIPLoss(device=self.device, seed=0, tokenizer=tokenize)
        state = torch.load(b32, map_location="cpu")
        state = state.state_dict() if hasattr(state, "state_dict") else state.get("state_dict", state)
        missing, _ = loss.model.load_openai_state_dict(state)
        if missing:
            raise RuntimeError(f"{b32} does not hold an OpenAI CLIP ViT-B/32 state dict: missing {list(missing)[:5]}")
        loss.model.to(self.device).eval().requires_grad_(False)
        return loss

    # -- stage 1: text (host) ---------------------------------------------------------------------------
    def encode(self, items: Sequence[WorkItem], pin: bool = True) -> Dict[str, torch.Tensor]:
        """Host-side conditioning of a batch of work items (all with the same number of objects):
        uc, c [B,77,768], locals [n_obj,B,77,768], x_T [B,4,h,w] — pinned host tensors."""
        B = len(items)
        n_obj = len(items[0].object_names)
        assert all(len(it.object_names) == n_obj for it in items)
        uc = self.text.encode([""] * B).cpu()
        c = self.text.encode([it.prompt for it in items]).cpu()
        loc = torch.stack([self.text.encode(["a photo of " + it.object_names[i] for it in items]).cpu()
                           for i in range(n_obj)]) if n_obj else torch.zeros(0, B, 77, 768)
        xs = []
        for it in items:
            g = torch.Generator().manual_seed(it.seed)
            xs.append(torch.randn(1, 4, self.latent_size, self.latent_size, generator=g))
        out = {"uc": uc, "c": c, "locals": loc, "x_T": torch.cat(xs)}
        if pin and torch.cuda.is_available():
            out = {k: v.contiguous().pin_memory() for k, v in out.items()}
        return out

    def to_device(self, cond: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {k: v.to(self.device, non_blocking=True) for k, v in cond.items()}

    @staticmethod
    def h2d_bytes(cond: Dict[str, torch.Tensor]) -> int:
        return sum(v.numel() * v.element_size() for v in cond.values())

    # -- stage 2: sampling + alpha optimisation (device) ----------------------------------------------
    def generate(self, items: Sequence[WorkItem], cond: Dict[str, torch.Tensor], optimize_alpha: bool = True,
                 alpha=None, to_host: bool = False, check_device_error: bool = True):
        """Run the sampler on a batch of work items.  `cond` may live on the host (pinned) or on the device.
        Returns the decoded images [B,3,H,W] in [0,1] (on the host when to_host=True) or, without a VAE, the latents."""
        if cond["c"].device != self.device:
            cond = self.to_device(cond)
        B = len(items)
        n_obj = len(items[0].object_names)
        with torch.autocast("cuda", dtype=torch.float16):
            self.sampler.sample(
                S=self.steps, batch_size=B, shape=[4, self.latent_size, self.latent_size], conditioning=cond["c"],
                x_T=cond["x_T"], unconditional_guidance_scale=self.scale, unconditional_conditioning=cond["uc"],
                eta=0.0, text_index=0, curr_text=[it.prompt for it in items],
                bboxs_curr=[it.bboxes for it in items] if B > 1 else items[0].bboxes,
                seed=items[0].seed, prompt_idx=[it.prompt_idx for it in items],
                object_names=[it.object_names for it in items] if B > 1 else items[0].object_names,
                local_conditionings=[cond["locals"][i] for i in range(n_obj)], optimize_alpha=optimize_alpha,
                alpha=alpha)
        res = self.sampler.last_result
        out = res["image"] if res["image"] is not None else res["latent"]
        if check_device_error:
            # every mbarrier wait inside the tcgen05 kernels is bounded; a timeout only sets a device-side word (CUDA-graph
            # replays have no per-launch return code), so it is read here, once per image (one sync per ~second of work)
            from . import native

            err = native.device_error()
            if err:
                raise RuntimeError(f"sta_b200 kernels reported device error 0x{err:x} (mbarrier wait timed out, id {err & 0xff}) "
                                   "while generating this image: results are invalid")
        if to_host:
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=False)
            return host
        return out
