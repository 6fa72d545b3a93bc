"""Host-side owner of one ``ug_ctx``: weights in, tensors through the C ABI, tensors out.

PyTorch is used for device memory and streams only; every FLOP of the UNet / VAE runs in
libunigeo_b200.so.  Tensors at this level are fp32 CUDA tensors in the upstream layouts
(see include/unigeo_b200.h), so callers read like the [UPSTREAM] pipeline they replace
(reference call site: /root/reference/model/depthcrafter.py:80-90).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Tuple

import torch

from . import _lib
from .config import PipelineConfig

_DTYPES = {"fp16": _lib.UG_F16, "float16": _lib.UG_F16, "bf16": _lib.UG_BF16, "bfloat16": _lib.UG_BF16}
_TORCH_TO_UG = {torch.float16: _lib.UG_F16, torch.bfloat16: _lib.UG_BF16, torch.float32: _lib.UG_F32}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


class Engine:
    """One context = one GPU = one model replica (clip-sharded data parallelism above it)."""

    def __init__(self, cfg: PipelineConfig, dtype: str = "fp16", device: int = 0, sn_cfg=None,
                 vae_encode_dtype: str | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("unigeo_b200.Engine needs a CUDA device (B200, sm_100a); there is no CPU path")
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        self.lib = _lib.load()
        self.cfg = cfg
        self.dtype = dtype
        self.device = torch.device("cuda", device)
        self._cfg_struct = _lib.cfg_struct(cfg, _DTYPES[dtype])
        self._ctx = C.c_void_p()
        _lib.check(self.lib.ug_ctx_create(C.byref(self._ctx), device, C.byref(self._cfg_struct)))
        if vae_encode_dtype is not None:   # the encoder alone in bf16 (upstream upcasts it to fp32: fp16 range)
            if vae_encode_dtype not in _DTYPES:
                raise ValueError(f"vae_encode_dtype must be one of {sorted(_DTYPES)}")
            _lib.check(self.lib.ug_ctx_set_vae_encode_dtype(self._ctx, _DTYPES[vae_encode_dtype]))
        self.vae_encode_dtype = vae_encode_dtype or dtype
        self._shape: Tuple[int, int, int] | None = None
        self._finalized = False
        self.sn_cfg = sn_cfg
        cc = getattr(cfg, "clip", None)
        if cc is not None:           # CLIP image encoder dims (weights optional; used by clip_embed)
            self._cfg_clip = _lib.ClipCfg(cc.hidden_size, cc.num_hidden_layers, cc.num_attention_heads,
                                          cc.intermediate_size, cc.patch_size, cc.image_size, cc.projection_dim,
                                          cc.layer_norm_eps)
            _lib.check(self.lib.ug_ctx_set_clip_cfg(self._ctx, C.byref(self._cfg_clip)))
        if sn_cfg is not None:       # StableNormal path: 2-D UNet / ControlNet dims (the 2-D VAE uses cfg.vae)
            if sn_cfg.unet2d.norm_groups != cfg.unet.norm_groups or sn_cfg.vae2d != cfg.vae:
                raise ValueError("the 2-D path shares GroupNorm groups and VAE dims with the pipeline config")
            self._cfg2d = _lib.unet2d_cfg_struct(sn_cfg)
            _lib.check(self.lib.ug_ctx_set_unet2d_cfg(self._ctx, C.byref(self._cfg2d)))

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, prefix: str, sd: Dict[str, torch.Tensor]) -> None:
        """``prefix`` is "unet" or "vae"; ``sd`` maps diffusers keys to tensors (any device/dtype)."""
        with torch.cuda.device(self.device):
            keep, keys = [], []
            for key, t in sd.items():
                if t.dtype not in _TORCH_TO_UG:
                    t = t.float()
                if prefix == "clip":             # see include/unigeo_b200.h: ug_clip_embed
                    if key.endswith("patch_embedding.weight"):
                        t = t.reshape(t.shape[0], -1)
                    elif key.endswith("position_embedding.weight"):
                        t = t.reshape(-1)
                keep.append(t.to(self.device).contiguous())
                keys.append(f"{prefix}.{key}".encode())
            # the whole dict in slices of <= 256 MB of staged source tensors: ONE conversion launch per slice
            i0 = 0
            while i0 < len(keep):
                i1, nbytes = i0, 0
                while i1 < len(keep) and (i1 == i0 or nbytes + keep[i1].numel() * keep[i1].element_size() <= (1 << 28)):
                    nbytes += keep[i1].numel() * keep[i1].element_size()
                    i1 += 1
                n = i1 - i0
                shapes = (C.c_int64 * (5 * n))()
                for j, t in enumerate(keep[i0:i1]):
                    for k, d in enumerate(t.shape):
                        shapes[5 * j + k] = d
                _lib.check(self.lib.ug_ctx_load_weights(
                    self._ctx, n, (C.c_char_p * n)(*keys[i0:i1]), (C.c_void_p * n)(*[t.data_ptr() for t in keep[i0:i1]]),
                    (C.c_int * n)(*[_TORCH_TO_UG[t.dtype] for t in keep[i0:i1]]), shapes,
                    (C.c_int * n)(*[t.dim() for t in keep[i0:i1]]), _stream()))
                i0 = i1
            torch.cuda.current_stream().synchronize()
        self._finalized = False

    def finalize(self) -> None:
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_ctx_finalize(self._ctx, _stream()))
        self._finalized = True

    def prepare(self, T: int, h: int, w: int) -> None:
        if not self._finalized:
            self.finalize()
        if self._shape != (T, h, w):
            with torch.cuda.device(self.device):
                _lib.check(self.lib.ug_ctx_prepare(self._ctx, T, h, w, _stream()))
            self._shape = (T, h, w)

    # ------------------------------------------------------------------ UNet / denoising
    def set_clip_context(self, enc: torch.Tensor) -> None:
        """enc: [T, cross_attention_dim] CLIP image embeddings of the clip's frames."""
        enc = _f32(enc, self.device)
        assert self._shape is not None and enc.shape == (self._shape[0], self.cfg.unet.cross_attention_dim)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_set_clip_context(self._ctx, enc.data_ptr(), _stream()))

    def unet_forward(self, x: torch.Tensor, timestep: float, added_time_ids: Iterable[float]) -> torch.Tensor:
        """x [1,T,8,h,w] -> v [1,T,4,h,w] (needs prepare + set_clip_context)."""
        x = _f32(x, self.device)
        _, T, _, h, w = x.shape
        assert self._shape == (T, h, w), "call prepare(T, h, w) first"
        ids = (C.c_float * 3)(*[float(v) for v in added_time_ids])
        out = torch.empty((1, T, self.cfg.unet.out_channels, h, w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_unet_st_forward(self._ctx, x.data_ptr(), float(timestep), ids, out.data_ptr(),
                                                   _stream()))
        return out

    def denoise(self, cond_latents: torch.Tensor, init_noise: torch.Tensor, added_time_ids, steps: int,
                out: torch.Tensor | None = None) -> torch.Tensor:
        """Whole Euler/Karras loop. cond_latents, init_noise: [T,4,h,w] fp32 -> latents [T,4,h,w]."""
        cond = _f32(cond_latents, self.device)
        noise = _f32(init_noise, self.device)
        T, _, h, w = cond.shape
        assert self._shape == (T, h, w), "call prepare(T, h, w) first"
        ids = (C.c_float * 3)(*[float(v) for v in added_time_ids])
        if out is None:
            out = torch.empty_like(cond)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_denoise_clip(self._ctx, cond.data_ptr(), noise.data_ptr(), ids, int(steps),
                                                out.data_ptr(), _stream()))
        return out

    # ------------------------------------------------------------------ VAE
    def vae_encode(self, img: torch.Tensor, noise: torch.Tensor | None = None, noise_strength: float = 0.0,
                   out: torch.Tensor | None = None) -> torch.Tensor:
        """img [N,3,H,W] in [-1,1] (+ noise * strength) -> latent mean [N,4,H/8,W/8]."""
        if not self._finalized:
            self.finalize()
        img = _f32(img, self.device)
        N, _, H, W = img.shape
        nptr = None
        if noise is not None:
            noise = _f32(noise, self.device)
            nptr = noise.data_ptr()
        if out is None:
            out = torch.empty((N, self.cfg.vae.latent_channels, H // 8, W // 8), dtype=torch.float32,
                              device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_vae_encode(self._ctx, img.data_ptr(), nptr, float(noise_strength), N, H, W,
                                              out.data_ptr(), _stream()))
        return out

    def vae_decode(self, latents: torch.Tensor, chunk: int = 8, out: torch.Tensor | None = None) -> torch.Tensor:
        """latents [T,4,h,w] (scaled) -> frames [T,3,8h,8w]."""
        if not self._finalized:
            self.finalize()
        lat = _f32(latents, self.device)
        T, _, h, w = lat.shape
        if out is None:
            out = torch.empty((T, self.cfg.vae.in_channels, 8 * h, 8 * w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_vae_decode_temporal(self._ctx, lat.data_ptr(), T, h, w, int(chunk),
                                                       out.data_ptr(), _stream()))
        return out

    def clip_embed(self, video: torch.Tensor) -> torch.Tensor:
        """video [F,3,H,W] in [-1,1] -> CLIP image embeddings [F, projection_dim] ("clip" weights)."""
        if not self._finalized:
            self.finalize()
        v = _f32(video, self.device)
        F_, _, H, W = v.shape
        out = torch.empty((F_, self.cfg.clip.projection_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_clip_embed(self._ctx, v.data_ptr(), F_, H, W, out.data_ptr(), _stream()))
        return out

    def prepare_frames(self, images: torch.Tensor) -> torch.Tensor:
        """images [T,3,H,W] fp32 0..255 (CUDA) -> frames [T,H,W,3] = float(uint8(v))/255 (reference :39-45)."""
        img = _f32(images, self.device)
        T, _, H, W = img.shape
        out = torch.empty((T, H, W, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_prepare_frames(self._ctx, img.data_ptr(), T, H, W, out.data_ptr(), _stream()))
        return out

    def vae_encode_frames(self, frames: torch.Tensor, noise: torch.Tensor | None = None, noise_strength: float = 0.0,
                          want_video: bool = False):
        """frames [T,H,W,3] fp32 in [0,1] (the adapter's ``prepare_input`` layout) -> latent mean [T,4,H/8,W/8];
        x*2-1 and the noise augmentation (noise [T,3,H,W]) are fused into the layout kernel.  With
        ``want_video`` also returns x*2-1 as [T,3,H,W] (input of the CLIP branch)."""
        if not self._finalized:
            self.finalize()
        fr = _f32(frames, self.device)
        T, H, W, _ = fr.shape
        nz = _f32(noise, self.device) if noise is not None else None
        out = torch.empty((T, self.cfg.vae.latent_channels, H // 8, W // 8), dtype=torch.float32, device=self.device)
        video = torch.empty((T, 3, H, W), dtype=torch.float32, device=self.device) if want_video else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_vae_encode_frames(self._ctx, fr.data_ptr(), nz.data_ptr() if nz is not None else None,
                                                     float(noise_strength), T, H, W,
                                                     video.data_ptr() if video is not None else None, out.data_ptr(),
                                                     _stream()))
        return (out, video) if want_video else out

    def vae_decode_frames(self, latents: torch.Tensor, chunk: int = 8) -> torch.Tensor:
        """latents [T,4,h,w] (scaled) -> ``.frames[0]``: [T,8h,8w,3] fp32 = clamp(decode/2+0.5, 0, 1)."""
        if not self._finalized:
            self.finalize()
        lat = _f32(latents, self.device)
        T, _, h, w = lat.shape
        out = torch.empty((T, 8 * h, 8 * w, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_vae_decode_frames(self._ctx, lat.data_ptr(), T, h, w, int(chunk), out.data_ptr(),
                                                     _stream()))
        return out

    def depth_postprocess(self, frames: torch.Tensor, intrinsics: torch.Tensor):
        """frames [T,H,W,3] in [0,1], intrinsics [T,3,3] -> (pred_depths [T,H,W], pred_normals [T,H,W,3] OpenGL)."""
        fr = _f32(frames, self.device)
        K = _f32(intrinsics, self.device)
        T, H, W, _ = fr.shape
        assert K.shape == (T, 3, 3)
        depth = torch.empty((T, H, W), dtype=torch.float32, device=self.device)
        normals = torch.empty((T, H, W, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_depth_postprocess(self._ctx, fr.data_ptr(), K.data_ptr(), T, H, W, depth.data_ptr(),
                                                     normals.data_ptr(), _stream()))
        return depth, normals

    # ------------------------------------------------------------------ metric kernels (eval.py:49 / :54)
    def _mask_u8(self, mask, n: int):
        if mask is None:
            return None
        m = torch.as_tensor(mask).to(self.device)
        m = (m != 0).to(torch.uint8).contiguous().reshape(-1)
        assert m.numel() == n, "custom_mask must have one entry per pixel"
        return m

    def depth_metrics(self, pred: torch.Tensor, gt: torch.Tensor, mask=None, max_depth: float = 80.0,
                      with_maps: bool = False):
        """Scale/shift-aligned depth metrics of one clip ([Nf,H,W] or [H,W] tensors, any device).
        Returns (list of 11 floats, maps | None): see ug_depth_metrics in include/unigeo_b200.h."""
        p, g = _f32(pred, self.device), _f32(gt, self.device)
        assert p.shape == g.shape
        n = p.numel()
        m = self._mask_u8(mask, n)
        out = (C.c_double * 11)()
        maps = [torch.empty_like(g) for _ in range(3)] if with_maps else [None] * 3
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_depth_metrics(self._ctx, p.data_ptr(), g.data_ptr(),
                                                 m.data_ptr() if m is not None else None, n, float(max_depth), out,
                                                 *[t.data_ptr() if t is not None else None for t in maps],
                                                 _stream()))
        return list(out), (tuple(maps) if with_maps else None)

    def normal_metrics(self, pred: torch.Tensor, gt: torch.Tensor, mask=None, with_map: bool = False):
        """Angular-error metrics of one clip ([Nf,H,W,3] tensors, any device) -> list of 8 floats
        (with_map: also the per-pixel error in degrees, [Nf,H,W])."""
        p, g = _f32(pred, self.device), _f32(gt, self.device)
        assert p.shape == g.shape and p.shape[-1] == 3
        n = p.numel() // 3
        m = self._mask_u8(mask, n)
        out = (C.c_double * 8)()
        err = torch.empty(p.shape[:-1], dtype=torch.float32, device=self.device) if with_map else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_normal_metrics(self._ctx, p.data_ptr(), g.data_ptr(),
                                                  m.data_ptr() if m is not None else None, n, out,
                                                  err.data_ptr() if with_map else None, _stream()))
        return (list(out), err) if with_map else list(out)

    # ------------------------------------------------------------------ scene stitch (sharding.stitch_scene)
    def stitch_fit(self, gathered: torch.Tensor, num_clips: int, disparity: bool = True, offset: float = 0.1):
        """gathered [world, per_rank, 2, overlap, H, W] fp32 CUDA (all-gathered heads / tails) -> chain [num_clips, 2]
        float64 CUDA: (S, T) of every clip into clip 0's frame (ug_stitch_fit)."""
        g = _f32(gathered, self.device)
        world, per_rank = g.shape[0], g.shape[1]
        n = g[0, 0, 0].numel()
        chain = torch.empty((num_clips, 2), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_stitch_fit(self._ctx, g.data_ptr(), world, per_rank, int(num_clips), n,
                                              1 if disparity else 0, float(offset), chain.data_ptr(), _stream()))
        return chain

    def stitch_apply(self, clip: torch.Tensor, prev_tail, chain: torch.Tensor, k: int, overlap: int,
                     disparity: bool = True, offset: float = 0.1) -> torch.Tensor:
        """clip [T,H,W] (clip k of the scene), prev_tail [overlap,H,W] | None -> clip k in clip 0's frame, head ramped."""
        d = _f32(clip, self.device)
        frame = d[0].numel()
        pt = _f32(prev_tail, self.device) if prev_tail is not None else None
        out = torch.empty_like(d)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_stitch_apply(self._ctx, d.data_ptr(), d.numel(), pt.data_ptr() if pt is not None else None,
                                                overlap * frame, frame, int(overlap), chain.data_ptr(), int(k),
                                                1 if disparity else 0, float(offset), out.data_ptr(), _stream()))
        return out

    # ------------------------------------------------------------------ StableNormal path (2-D UNet)
    def set_text_context(self, net: str, tokens: torch.Tensor) -> None:
        """tokens [L,D] (shared prompt) or [F,L,D]: encoder_hidden_states of network ``net`` ("unet2d", ...)."""
        if not self._finalized:
            self.finalize()
        t = _f32(tokens, self.device)
        if t.dim() == 2:
            t = t.unsqueeze(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_set_text_context(self._ctx, net.encode(), t.data_ptr(), t.shape[0], t.shape[1],
                                                    _stream()))

    def unet2d_forward(self, net: str, x: torch.Tensor, timestep: float, controlnet: str | None = None,
                       controlnet_sample: torch.Tensor | None = None) -> torch.Tensor:
        """x [F,Cin,h,w] -> [F,Cout,h,w]; with ``controlnet`` its residuals (computed from
        ``controlnet_sample``) are added to the skips / mid block."""
        if not self._finalized:
            self.finalize()
        x = _f32(x, self.device)
        F_, _, h, w = x.shape
        out = torch.empty((F_, self.sn_cfg.unet2d.out_channels, h, w), dtype=torch.float32, device=self.device)
        cs = _f32(controlnet_sample, self.device) if controlnet is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_unet2d_forward(self._ctx, net.encode(), x.data_ptr(), F_, h, w, float(timestep),
                                                  controlnet.encode() if controlnet else None,
                                                  cs.data_ptr() if cs is not None else None, out.data_ptr(),
                                                  _stream()))
        return out

    def refine_2d(self, net: str, controlnet: str | None, image_latent: torch.Tensor | None,
                  latents: torch.Tensor, steps: int, t_start: int = -1) -> torch.Tensor:
        """DDIM ("sample" prediction) refinement loop over [F,4,h,w] latents."""
        if not self._finalized:
            self.finalize()
        lat = _f32(latents, self.device)
        F_, _, h, w = lat.shape
        il = _f32(image_latent, self.device) if image_latent is not None else None
        out = torch.empty_like(lat)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_refine_frames_2d(self._ctx, net.encode(), controlnet.encode() if controlnet else None,
                                                    il.data_ptr() if il is not None else None, lat.data_ptr(), F_, h, w,
                                                    int(steps), int(t_start), out.data_ptr(), _stream()))
        return out

    def vae2d_encode(self, img: torch.Tensor, out_scale: float = 1.0) -> torch.Tensor:
        """img [N,3,H,W] in [-1,1] -> latent mode * out_scale [N,4,H/8,W/8] ("vae2d" weights)."""
        if not self._finalized:
            self.finalize()
        img = _f32(img, self.device)
        N, _, H, W = img.shape
        out = torch.empty((N, self.cfg.vae.latent_channels, H // 8, W // 8), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_vae2d_encode(self._ctx, img.data_ptr(), N, H, W, float(out_scale), out.data_ptr(),
                                                _stream()))
        return out

    def vae2d_decode(self, latents: torch.Tensor, want_image: bool = True, want_normals_u8: bool = False):
        """latents [N,4,h,w] (scaled) -> (image [N,3,8h,8w] fp32 | None, unit normals uint8 [N,8h,8w,3] | None)."""
        if not self._finalized:
            self.finalize()
        lat = _f32(latents, self.device)
        N, _, h, w = lat.shape
        img = torch.empty((N, 3, 8 * h, 8 * w), dtype=torch.float32, device=self.device) if want_image else None
        nrm = torch.empty((N, 8 * h, 8 * w, 3), dtype=torch.uint8, device=self.device) if want_normals_u8 else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ug_vae2d_decode(self._ctx, lat.data_ptr(), N, h, w,
                                                img.data_ptr() if img is not None else None,
                                                nrm.data_ptr() if nrm is not None else None, _stream()))
        return img, nrm

    # ------------------------------------------------------------------ bookkeeping
    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.ug_ctx_launch_count(self._ctx, 1 if reset else 0))

    def profile(self, enable, by_shape: bool = False) -> None:
        _lib.check(self.lib.ug_ctx_profile(self._ctx, (2 if by_shape else 1) if enable else 0))

    def profile_read(self, cap: int = 512):
        """[{name, launches, ms, flops, bytes}] aggregated over the launches since profile(True)."""
        names = C.create_string_buffer(64 * cap)
        cnt = (C.c_longlong * cap)()
        ms, fl, by = (C.c_double * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
        n = self.lib.ug_ctx_profile_read(self._ctx, cap, names, cnt, ms, fl, by)
        if n < 0:
            _lib.check(n)
        rows = []
        for i in range(n):
            nm = names.raw[64 * i:64 * i + 64].split(b"\0", 1)[0].decode()
            rows.append(dict(name=nm, launches=int(cnt[i]), ms=float(ms[i]), flops=float(fl[i]), bytes=float(by[i])))
        return rows

    def graph_count(self) -> int:
        """CUDA graphs instantiated for the step loops (0 until a loop signature has been called twice)."""
        return int(self.lib.ug_ctx_graph_count(self._ctx))

    def workspace_bytes(self) -> int:
        return int(self.lib.ug_ctx_workspace_bytes(self._ctx))

    def close(self) -> None:
        if self._ctx:
            self.lib.ug_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
