"""Host-side mirror of the reference's module surface for the diffusion sampling path, bound to the
C ABI (libsdxl_b200.so). PyTorch is used only for device memory and streams.

Mirrored reference interfaces (file:line relative to the reference root):
  UNet::forward(x, timesteps, context, label)                 src/model/unet/mod.rs:449-493
  Diffuser::sample_latent / sample_latent_with_inpainting /
            refine_latent                                     src/model/stablediffusion/mod.rs:317-376
  Conditioning                                                src/model/stablediffusion/mod.rs:544-555
  Backend::qkv_attention                                      src/backend.rs:4-10
Error behaviour: the reference panics on shape errors; here every failure raises SdxlError carrying
sdxl_last_error().
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import SdxlError
from .config import UNetConfig, VaeConfig
from .weights import build_pack


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Context:
    """One (device, stream). Replaces the reference's fixed LibTorchDevice::Cuda(0) (sample/main.rs:131)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise SdxlError("sdxl_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(self.device)
        h = C.c_void_p()
        rc = self.lib.sdxl_ctx_create(device, C.c_void_p(self.stream.cuda_stream), C.byref(h))
        if rc != 0:
            raise SdxlError(f"sdxl_ctx_create failed with {rc}")
        self.h = h

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.lib.sdxl_last_error(self.h)
            raise SdxlError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def enter(self) -> None:
        """Order the ctx stream after work already queued on torch's current stream."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))

    def leave(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def synchronize(self) -> None:
        self.check(self.lib.sdxl_ctx_synchronize(self.h), "sdxl_ctx_synchronize")

    @property
    def launch_count(self) -> int:
        return int(self.lib.sdxl_ctx_launch_count(self.h))

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.sdxl_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- operator level ------------------------------------------------------------------------
    def qkv_attention(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: Optional[torch.Tensor],
                      n_head: int) -> torch.Tensor:
        """== Backend::qkv_attention (src/backend.rs:4-10). q [B,T,C], k/v [B,S,C] f16; mask: additive [T,S] (the text
        encoders' decoder mask) or None (UNet / VAE)."""
        B, T, Cc = q.shape
        S = k.shape[1]
        q, k, v = (t.to(self.device, torch.float16).contiguous() for t in (q, k, v))
        if mask is not None:
            if tuple(mask.shape) != (T, S):
                raise SdxlError(f"qkv_attention: mask must be [{T},{S}], got {tuple(mask.shape)}")
            mask = mask.to(self.device, torch.float16).contiguous()
        out = torch.empty_like(q)
        self.enter()
        self.check(self.lib.sdxl_qkv_attention(self.h, _ptr(q), _ptr(k), _ptr(v), _ptr(mask), B, T, S, Cc, n_head, _ptr(out)),
                   "sdxl_qkv_attention")
        self.leave()
        return out

    def linear(self, x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
               residual: Optional[torch.Tensor] = None, geglu: bool = False, out_f16: bool = False) -> torch.Tensor:
        """== nn::Linear::forward, weight [in,out]. x [M,K] f16."""
        x = x.to(self.device, torch.float16).contiguous()
        w = w.to(self.device, torch.float16).contiguous()
        M, K = x.shape
        N = w.shape[1]
        bias = None if bias is None else bias.to(self.device, torch.float16).contiguous()
        residual = None if residual is None else residual.to(self.device, torch.float32).contiguous()
        if geglu:
            out = torch.empty(M, N // 2, device=self.device, dtype=torch.float16)
        else:
            out = torch.empty(M, N, device=self.device, dtype=torch.float16 if out_f16 else torch.float32)
        self.enter()
        self.check(self.lib.sdxl_op_linear(self.h, _ptr(x), _ptr(w), _ptr(bias), _ptr(residual), M, K, N, int(geglu),
                                           int(out_f16), _ptr(out)), "sdxl_op_linear")
        self.leave()
        return out

    def conv2d(self, x_nhwc: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1,
               upsample: bool = False) -> torch.Tensor:
        """== Conv2d::forward on NHWC f32 input, OIHW f16 weight; pad = k//2."""
        x = x_nhwc.to(self.device, torch.float32).contiguous()
        w = w.to(self.device, torch.float16).contiguous()
        bias = None if bias is None else bias.to(self.device, torch.float16).contiguous()
        B, H, W, Cin = x.shape
        Cout, _, ks, _ = w.shape
        Ho, Wo = (H // 2, W // 2) if stride == 2 else ((2 * H, 2 * W) if upsample else (H, W))
        out = torch.empty(B, Ho, Wo, Cout, device=self.device, dtype=torch.float32)
        self.enter()
        self.check(self.lib.sdxl_op_conv2d(self.h, _ptr(x), _ptr(w), _ptr(bias), B, H, W, Cin, Cout, ks, stride,
                                           int(upsample), _ptr(out)), "sdxl_op_conv2d")
        self.leave()
        return out

    def group_norm(self, x1: torch.Tensor, x2: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                   n_group: int = 32, eps: float = 1e-5, silu: bool = False) -> torch.Tensor:
        """== GroupNorm::forward (+SiLU) on NHWC f32 [B,HW,C]; x2 is channel-concatenated after x1."""
        x1 = x1.to(self.device, torch.float32).contiguous()
        x2 = None if x2 is None else x2.to(self.device, torch.float32).contiguous()
        B, HW, C1 = x1.shape
        C2 = 0 if x2 is None else x2.shape[2]
        gamma = gamma.to(self.device, torch.float32).contiguous()
        beta = beta.to(self.device, torch.float32).contiguous()
        out = torch.empty(B, HW, C1 + C2, device=self.device, dtype=torch.float16)
        self.enter()
        self.check(self.lib.sdxl_op_group_norm(self.h, _ptr(x1), C1, _ptr(x2), C2, B, HW, n_group, _ptr(gamma), _ptr(beta),
                                               eps, int(silu), _ptr(out)), "sdxl_op_group_norm")
        self.leave()
        return out

    def layer_norm(self, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
        x = x.to(self.device, torch.float32).contiguous()
        rows, Cc = x.shape
        gamma = gamma.to(self.device, torch.float32).contiguous()
        beta = beta.to(self.device, torch.float32).contiguous()
        out = torch.empty(rows, Cc, device=self.device, dtype=torch.float16)
        self.enter()
        self.check(self.lib.sdxl_op_layer_norm(self.h, _ptr(x), _ptr(gamma), _ptr(beta), eps, rows, Cc, _ptr(out)),
                   "sdxl_op_layer_norm")
        self.leave()
        return out

    def timestep_embedding(self, timesteps: Sequence[int], dim: int, max_period: int = 10000) -> torch.Tensor:
        n = len(timesteps)
        arr = (C.c_int32 * n)(*[int(t) for t in timesteps])
        out = torch.empty(n, dim, device=self.device, dtype=torch.float32)
        self.enter()
        self.check(self.lib.sdxl_op_timestep_embedding(self.h, arr, n, dim, max_period, _ptr(out)),
                   "sdxl_op_timestep_embedding")
        self.leave()
        return out

    def randn(self, n: int, seed: int, subsequence: int = 0) -> torch.Tensor:
        out = torch.empty(n, device=self.device, dtype=torch.float32)
        self.enter()
        self.check(self.lib.sdxl_randn(self.h, _ptr(out), n, seed, subsequence), "sdxl_randn")
        self.leave()
        return out


@dataclass
class Conditioning:
    """Mirror of the reference's Conditioning record (stablediffusion/mod.rs:544-555); f16 tensors."""
    context_full: Optional[torch.Tensor] = None                       # [B,77,2048]
    context_open_clip: Optional[torch.Tensor] = None                  # [B,77,1280]
    unconditional_context_full: Optional[torch.Tensor] = None         # [77,2048]
    unconditional_context_open_clip: Optional[torch.Tensor] = None    # [77,1280]
    channel_context: Optional[torch.Tensor] = None                    # [B,2816]
    channel_context_refiner: Optional[torch.Tensor] = None            # [B,2560]
    unconditional_channel_context: Optional[torch.Tensor] = None      # [2816]
    unconditional_channel_context_refiner: Optional[torch.Tensor] = None  # [2560]
    resolution: Sequence[int] = (1024, 1024)                          # (height, width)

    def _fields(self) -> List[str]:
        return ["context_full", "context_open_clip", "unconditional_context_full", "unconditional_context_open_clip",
                "channel_context", "channel_context_refiner", "unconditional_channel_context",
                "unconditional_channel_context_refiner"]

    def to_struct(self, device: Optional[torch.device]):
        """Returns (ctypes struct, keep-alive list). device=None => host pointers (pinned f16 copies)."""
        keep = []
        s = _lib.Conditioning()
        on_host = device is None
        s.on_host = int(on_host)
        ref = self.context_full if self.context_full is not None else self.context_open_clip
        s.n_batch = int(ref.shape[0])
        s.n_ctx = int(ref.shape[1])
        for f in self._fields():
            t = getattr(self, f)
            if t is None:
                setattr(s, f, None)
                continue
            t = t.to(torch.float16)
            t = t.cpu().contiguous() if on_host else t.to(device).contiguous()
            keep.append(t)
            setattr(s, f, t.data_ptr())
        s.resolution[0], s.resolution[1] = int(self.resolution[0]), int(self.resolution[1])
        return s, keep


def _cfg_struct(cfg: UNetConfig) -> _lib.UnetCfg:
    s = _lib.UnetCfg()
    s.adm_in_channels = cfg.adm_in_channels
    s.in_channels = cfg.in_channels
    s.out_channels = cfg.out_channels
    s.model_channels = cfg.model_channels
    s.n_levels = cfg.n_levels
    for i, m in enumerate(cfg.channel_mults):
        s.channel_mults[i] = m
    for i, d in enumerate(cfg.transformer_depths):
        s.transformer_depths[i] = d
    s.n_head_channels = cfg.n_head_channels
    s.context_dim = cfg.context_dim
    s.is_refiner = int(cfg.is_refiner)
    s.n_steps = cfg.n_steps
    return s


class Diffuser:
    """Mirror of the reference's Diffuser (UNet + alphas + sampler loops), device-resident."""

    def __init__(self, ctx: Context, cfg: UNetConfig, weights, nccl_comm=None, rank: int = 0, root: int = 0):
        """weights: dict name->f16 tensor (reference layouts) or an already-built pack (uint8 tensor).
        nccl_comm (an ncclComm_t, see sharding.nccl_comm_init): multi-GPU load through sdxl_unet_load_broadcast — only `root`
        passes weights, every other rank passes None and receives the pack over NCCL inside the library."""
        self.ctx, self.cfg = ctx, cfg
        h = C.c_void_p()
        cs = _cfg_struct(cfg)
        pack = None
        if weights is not None:
            pack = weights if isinstance(weights, torch.Tensor) else build_pack(weights)
        on_device = bool(pack is not None and pack.is_cuda)
        ctx.enter()
        if on_device:
            torch.cuda.current_stream(ctx.device).synchronize()
        if nccl_comm is not None:
            rc = ctx.lib.sdxl_unet_load_broadcast(ctx.h, C.byref(cs), None if pack is None else pack.data_ptr(),
                                                  0 if pack is None else pack.numel(), int(on_device), nccl_comm, rank, root, C.byref(h))
            ctx.check(rc, "sdxl_unet_load_broadcast")
        else:
            rc = ctx.lib.sdxl_unet_load(ctx.h, C.byref(cs), pack.data_ptr(), pack.numel(), int(on_device), C.byref(h))
            ctx.check(rc, "sdxl_unet_load")
        self.h = h
        self._cond_key = None
        self._keep = None

    def close(self) -> None:
        if getattr(self, "h", None):
            self.ctx.lib.sdxl_unet_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- UNet::forward -------------------------------------------------------------------------
    def set_conditioning(self, context: torch.Tensor, label: torch.Tensor) -> None:
        ctx = self.ctx
        context = context.to(ctx.device, torch.float16).contiguous()
        label = label.to(ctx.device, torch.float16).contiguous()
        B, n_ctx, _ = context.shape
        ctx.enter()
        ctx.check(ctx.lib.sdxl_unet_set_conditioning(self.h, B, n_ctx, _ptr(context), _ptr(label)),
                  "sdxl_unet_set_conditioning")
        ctx.leave()
        self._keep = (context, label)

    def unet_forward(self, x: torch.Tensor, timesteps, context: Optional[torch.Tensor] = None,
                     label: Optional[torch.Tensor] = None) -> torch.Tensor:
        """== UNet::forward(x [B,4,h,w], timesteps Int[1], context [B,77,Cctx], label [B,adm]).
        f32 in -> f32 out (no I/O rounding), f16 in -> f16 out (the reference's tensors)."""
        ctx = self.ctx
        if context is not None:
            self.set_conditioning(context, label)
        t = int(timesteps[0]) if hasattr(timesteps, "__len__") else int(timesteps)
        B, _, h, w = x.shape
        # convert first, then enter: the ctx stream must wait for the cast / copy kernels torch queues on its own stream
        f16 = x.dtype == torch.float16
        x = x.to(ctx.device, torch.float16 if f16 else torch.float32).contiguous()
        out = torch.empty_like(x)
        ctx.enter()
        if f16:
            rc = ctx.lib.sdxl_unet_forward(self.h, B, h, w, _ptr(x), t, _ptr(out))
        else:
            rc = ctx.lib.sdxl_unet_forward_f32(self.h, B, h, w, _ptr(x), t, _ptr(out))
        ctx.check(rc, "sdxl_unet_forward")
        ctx.leave()
        return out

    @property
    def plan_flops(self) -> float:
        return float(self.ctx.lib.sdxl_unet_plan_flops(self.h))

    @property
    def plan_flops_executed(self) -> float:
        return float(self.ctx.lib.sdxl_unet_plan_flops_executed(self.h))

    @property
    def plan_num_ops(self) -> int:
        return int(self.ctx.lib.sdxl_unet_plan_num_ops(self.h))

    KIND_NAMES = ["igemm_tcgen05", "attention_tcgen05", "group_norm", "layer_norm", "gemv", "timestep_embedding",
                  "conv_in", "upsample2x", "phase_split", "cast_f16"]

    def profile_plan(self) -> Dict[str, Dict[str, float]]:
        """Per-kernel-kind device time (ms), algorithmic FLOPs and launch count of one plan execution."""
        ms, fl, ln = (C.c_double * _lib.PROFILE_KINDS)(), (C.c_double * _lib.PROFILE_KINDS)(), (C.c_int * _lib.PROFILE_KINDS)()
        self.ctx.check(self.ctx.lib.sdxl_unet_profile_plan(self.h, ms, fl, ln), "sdxl_unet_profile_plan")
        return {n: {"ms": ms[i], "flops": fl[i], "launches": ln[i]} for i, n in enumerate(self.KIND_NAMES) if ln[i]}

    def profile_dump(self, path: str) -> None:
        self.ctx.check(self.ctx.lib.sdxl_unet_profile_dump(self.h, path.encode()), "sdxl_unet_profile_dump")

    def alpha(self, i: int) -> float:
        return float(self.ctx.lib.sdxl_unet_alpha(self.h, i))

    # ---- Diffuser::* ---------------------------------------------------------------------------
    def _sample(self, cond: Conditioning, guidance: float, n_steps: int, step_start: int,
                init_latent: Optional[torch.Tensor], noise: Optional[torch.Tensor], seed: int,
                ref: Optional[torch.Tensor], mask: Optional[torch.Tensor], host: bool = False) -> torch.Tensor:
        ctx = self.ctx
        s, keep = cond.to_struct(None if host else ctx.device)
        h, w = cond.resolution[0] // 8, cond.resolution[1] // 8
        dev = torch.device("cpu") if host else ctx.device

        def prep(t, dt):
            return None if t is None else t.to(dev, dt).contiguous()
        init_latent = prep(init_latent, torch.float32)
        noise = prep(noise, torch.float32)
        ref = prep(ref, torch.float32)
        mask = prep(mask, torch.uint8)
        n_noise = 0 if noise is None else (noise.shape[0] if noise.dim() == 5 else 1)
        out = torch.empty(s.n_batch, self.cfg.in_channels, h, w, device=dev, dtype=torch.float32)
        ctx.enter()
        rc = ctx.lib.sdxl_sample_latent(self.h, C.byref(s), float(guidance), n_steps, step_start, _ptr(init_latent),
                                        _ptr(noise), n_noise, seed, _ptr(ref), _ptr(mask), _ptr(out))
        ctx.check(rc, "sdxl_sample_latent")
        ctx.leave()
        del keep
        return out

    def sample_latent(self, conditioning: Conditioning, unconditional_guidance_scale: float, n_steps: int,
                      noise: Optional[torch.Tensor] = None, seed: int = 0, host: bool = False) -> torch.Tensor:
        """== Diffuser::sample_latent (mod.rs:317-332). `noise` injects gen_noise()'s tensor (the reference's
        RNG is unseeded libtorch Philox; parity tests inject it)."""
        return self._sample(conditioning, unconditional_guidance_scale, n_steps, 0, noise, None, seed, None, None, host)

    def sample_latent_with_inpainting(self, conditioning: Conditioning, unconditional_guidance_scale: float,
                                      n_steps: int, reference: torch.Tensor, mask: torch.Tensor,
                                      init_noise: Optional[torch.Tensor] = None,
                                      step_noise: Optional[torch.Tensor] = None, seed: int = 0) -> torch.Tensor:
        """== Diffuser::sample_latent_with_inpainting (mod.rs:334-353); mask True keeps the generated latent."""
        return self._sample(conditioning, unconditional_guidance_scale, n_steps, 0, init_noise, step_noise, seed,
                            reference, mask.to(torch.uint8))

    def refine_latent(self, latent: torch.Tensor, conditioning: Conditioning, unconditional_guidance_scale: float,
                      step_start: int, n_steps: int, noise: Optional[torch.Tensor] = None, seed: int = 0) -> torch.Tensor:
        """== Diffuser::refine_latent (mod.rs:355-376)."""
        return self._sample(conditioning, unconditional_guidance_scale, n_steps, step_start, latent, noise, seed, None, None)

    # ---- step-wise (bench) ---------------------------------------------------------------------
    def sampler_begin(self, cond: Conditioning, guidance: float) -> None:
        s, keep = cond.to_struct(self.ctx.device)
        self.ctx.enter()
        self.ctx.check(self.ctx.lib.sdxl_sampler_begin(self.h, C.byref(s), float(guidance)), "sdxl_sampler_begin")
        self.ctx.synchronize()
        del keep

    def sampler_set_latent(self, x: torch.Tensor) -> None:
        x = x.to(self.ctx.device, torch.float32).contiguous()
        self.ctx.enter()
        self.ctx.check(self.ctx.lib.sdxl_sampler_set_latent(self.h, _ptr(x), 0), "sdxl_sampler_set_latent")
        self.ctx.synchronize()

    def sampler_get_latent(self, like: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(like, dtype=torch.float32, device=self.ctx.device)
        self.ctx.check(self.ctx.lib.sdxl_sampler_get_latent(self.h, _ptr(out), 0), "sdxl_sampler_get_latent")
        self.ctx.synchronize()
        return out

    def sampler_step(self, t: int, t_prev: int) -> None:
        self.ctx.check(self.ctx.lib.sdxl_sampler_step(self.h, t, t_prev), "sdxl_sampler_step")

    def sampler_step_host(self, t: int, t_prev: int, latent_host: torch.Tensor) -> None:
        assert latent_host.device.type == "cpu" and latent_host.dtype == torch.float32 and latent_host.is_contiguous()
        self.ctx.check(self.ctx.lib.sdxl_sampler_step_host(self.h, t, t_prev, latent_host.data_ptr()),
                       "sdxl_sampler_step_host")


class LatentDecoder:
    """Mirror of the reference's LatentDecoder (decode half): `decode_latent` and `latent_to_image`
    (src/model/stablediffusion/mod.rs:199-237, 263-266) over the device-resident VAE decoder."""

    KIND_NAMES = Diffuser.KIND_NAMES + ["softmax_rows", "transpose_f16", "post_quant"]

    def __init__(self, ctx: Context, cfg: VaeConfig, weights):
        self.ctx, self.cfg = ctx, cfg
        pack = weights if isinstance(weights, torch.Tensor) else build_pack(weights)
        on_device = pack.is_cuda
        ctx.enter()
        if on_device:
            torch.cuda.current_stream(ctx.device).synchronize()
        cs = _lib.VaeCfg()
        cs.latent_channels, cs.n_blocks, cs.n_group = cfg.latent_channels, len(cfg.block_channels), cfg.n_group
        cs.scale_factor = cfg.scale_factor
        for i, (ci, co) in enumerate(cfg.block_channels):
            cs.block_in[i], cs.block_out[i] = ci, co
        cs.n_enc_blocks, cs.enc_z_channels = len(cfg.enc_block_channels), cfg.enc_z_channels
        for i, (ci, co) in enumerate(cfg.enc_block_channels):
            cs.enc_in[i], cs.enc_out[i] = ci, co
        h = C.c_void_p()
        ctx.check(ctx.lib.sdxl_vae_load(ctx.h, C.byref(cs), pack.data_ptr(), pack.numel(), int(on_device), C.byref(h)),
                  "sdxl_vae_load")
        self.h = h

    def close(self) -> None:
        if getattr(self, "h", None):
            self.ctx.lib.sdxl_vae_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _prep(self, latent: torch.Tensor):
        host = not latent.is_cuda
        latent = latent.to(torch.float32).contiguous()
        B, _, h, w = latent.shape
        up = self.cfg.upscale
        return latent, host, B, h, w, up

    def decode_latent(self, latent: torch.Tensor) -> torch.Tensor:
        """latent f32 [B,4,h,w] (host or device) -> image f32 [B,3,8h,8w] on the same side."""
        latent, host, B, h, w, up = self._prep(latent)
        out = torch.empty((B, 3, h * up, w * up), dtype=torch.float32, device="cpu" if host else self.ctx.device)
        self.ctx.enter()
        self.ctx.check(self.ctx.lib.sdxl_vae_decode_latent(self.h, B, h, w, _ptr(latent), int(host), _ptr(out)),
                       "sdxl_vae_decode_latent")
        self.ctx.leave()
        return out

    def latent_to_image(self, latent: torch.Tensor) -> torch.Tensor:
        """RawImages buffer: u8 [B, 8h, 8w, 3]."""
        latent, host, B, h, w, up = self._prep(latent)
        out = torch.empty((B, h * up, w * up, 3), dtype=torch.uint8, device="cpu" if host else self.ctx.device)
        self.ctx.enter()
        self.ctx.check(self.ctx.lib.sdxl_vae_latent_to_image(self.h, B, h, w, _ptr(latent), int(host), _ptr(out)),
                       "sdxl_vae_latent_to_image")
        self.ctx.leave()
        return out

    def encode_image(self, image: torch.Tensor) -> torch.Tensor:
        """== LatentDecoder::encode_image: image f32 [B,3,H,W] in [-1,1] (host or device) -> latent f32 [B,4,H/8,W/8]."""
        host = not image.is_cuda
        image = image.to(torch.float32).contiguous()
        B, _, H, W = image.shape
        d = 2 ** (len(self.cfg.enc_block_channels) - 1)
        out = torch.empty((B, self.cfg.latent_channels, H // d, W // d), dtype=torch.float32, device="cpu" if host else self.ctx.device)
        self.ctx.enter()
        self.ctx.check(self.ctx.lib.sdxl_vae_encode_image(self.h, B, H, W, _ptr(image), int(host), _ptr(out)), "sdxl_vae_encode_image")
        self.ctx.leave()
        return out

    def image_to_latent(self, rgb: torch.Tensor) -> torch.Tensor:
        """== LatentDecoder::image_to_latent: RawImages u8 [B,H,W,3] -> latent f32 [B,4,H/8,W/8]."""
        host = not rgb.is_cuda
        rgb = rgb.to(torch.uint8).contiguous()
        B, H, W, _ = rgb.shape
        d = 2 ** (len(self.cfg.enc_block_channels) - 1)
        out = torch.empty((B, self.cfg.latent_channels, H // d, W // d), dtype=torch.float32, device="cpu" if host else self.ctx.device)
        self.ctx.enter()
        self.ctx.check(self.ctx.lib.sdxl_vae_image_to_latent(self.h, B, H, W, _ptr(rgb), int(host), _ptr(out)), "sdxl_vae_image_to_latent")
        self.ctx.leave()
        return out

    @property
    def encode_plan_flops(self) -> float:
        return float(self.ctx.lib.sdxl_vae_encode_plan_flops(self.h))

    @property
    def plan_flops(self) -> float:
        return float(self.ctx.lib.sdxl_vae_plan_flops(self.h))

    def profile_plan(self) -> Dict[str, Dict[str, float]]:
        ms, fl, ln = (C.c_double * _lib.PROFILE_KINDS)(), (C.c_double * _lib.PROFILE_KINDS)(), (C.c_int * _lib.PROFILE_KINDS)()
        self.ctx.check(self.ctx.lib.sdxl_vae_profile_plan(self.h, ms, fl, ln), "sdxl_vae_profile_plan")
        return {n: {"ms": ms[i], "flops": fl[i], "launches": ln[i]} for i, n in enumerate(self.KIND_NAMES) if ln[i]}

    def profile_dump(self, path: str) -> None:
        self.ctx.check(self.ctx.lib.sdxl_vae_profile_dump(self.h, path.encode()), "sdxl_vae_profile_dump")


def ddim_timesteps(n_steps: int, step_start: int = 0, total: int = 1000) -> List[int]:
    """(0..total-step_start).rev().step_by(total / n_steps)  (reference mod.rs:400-406)."""
    step = total // n_steps
    return list(range(total - step_start - 1, -1, -step))
