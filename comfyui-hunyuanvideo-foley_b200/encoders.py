"""Condition encoders on the engine: ctypes binding of the foley_encoder_* entry points.

`SiglipVisionEncoder.encode(pixels)` replaces `deps.siglip2_model.get_image_features(pixel_values=...)`
(reference feature_utils.py:64-79) and `ClapTextEncoder.encode(ids, mask)` replaces
`deps.clap_model(**inputs, output_hidden_states=True).last_hidden_state` (feature_utils.py:132-138), both as the bf16
modules the reference runs after `hunyuan_deps[key].to(device, dtype=target_dtype)` (nodes.py:283-284).  Weights come
from the HF state dict (same key names) or straight from the snapshot's model.safetensors; the tokenizer stays HF's.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

import torch

from .engine import FOLEY_DT, FoleyError, _check, _stream_ptr, load_library

ENC_SIGLIP_VISION, ENC_CLAP_TEXT, ENC_SYNCHFORMER = 0, 1, 2


class _EncConfig(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("kind", "hidden_size", "num_heads", "num_layers", "intermediate_size")] + \
               [("layer_norm_eps", c_float)] + \
               [(n, c_int32) for n in ("image_size", "patch_size", "vocab_size", "max_positions", "pad_token_id",
                                        "max_frames_per_pass")]


# google/siglip2-base-patch16-512 (vision_config) and laion/larger_clap_general (text_config): the two checkpoints the
# reference's Dependencies Loader names (nodes.py:199-201)
SIGLIP2_BASE_512 = dict(hidden_size=768, num_heads=12, num_layers=12, intermediate_size=3072, layer_norm_eps=1e-6,
                        image_size=512, patch_size=16)
MOTIONFORMER_DIVIDED_224 = dict(hidden_size=768, num_heads=12, num_layers=12, intermediate_size=3072, layer_norm_eps=1e-6,
                                image_size=224, patch_size=16)      # models/synchformer/divided_224_16x4.yaml
CLAP_TEXT_GENERAL = dict(hidden_size=768, num_heads=12, num_layers=12, intermediate_size=3072, layer_norm_eps=1e-12,
                         vocab_size=50265, max_positions=514, pad_token_id=1)

_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        lib.foley_encoder_create.argtypes = [POINTER(_EncConfig), c_int32, POINTER(c_void_p)]
        lib.foley_encoder_destroy.argtypes = [c_void_p]
        lib.foley_encoder_destroy.restype = None
        lib.foley_encoder_load_tensor.argtypes = [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32, c_int32]
        lib.foley_encoder_load_safetensors.argtypes = [c_void_p, c_char_p, c_char_p, POINTER(c_int64)]
        lib.foley_encoder_finalize.argtypes = [c_void_p]
        lib.foley_siglip_encode.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]
        lib.foley_clap_text_encode.argtypes = [c_void_p, POINTER(c_int32), POINTER(c_int32), c_int32, c_int32, c_void_p, c_void_p]
        lib.foley_synchformer_encode.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]
        lib.foley_encoder_set_option.argtypes = [c_void_p, c_char_p, c_int64]
        lib.foley_encoder_launch_count.argtypes = [c_void_p]
        lib.foley_encoder_launch_count.restype = c_int64
        lib.foley_encoder_debug_read.argtypes = [c_void_p, c_char_p, c_void_p, c_int64, POINTER(c_int64)]
        lib.foley_attention_d64.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                            c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_int32,
                                            c_int32, c_void_p]
        _bound = True
    return lib


class _Cfg(dict):
    """dict with attribute access: `deps.clap_model.config.max_position_embeddings` (reference utils.py:97-101 `_caps`)."""
    __getattr__ = dict.get


class _Encoder:
    KIND = None
    DEFAULTS = {}

    def __init__(self, config=None, device=None, **overrides):
        self.lib = _lib()
        if not torch.cuda.is_available():
            raise FoleyError("foley_b200 needs a CUDA device (sm_100a); no CPU path exists")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        c = dict(self.DEFAULTS)
        c.update(config or {})
        c.update(overrides)
        self.config = _Cfg(c)
        if "max_positions" in c:
            self.config["max_position_embeddings"] = c["max_positions"]
        self.cfg = _EncConfig(kind=self.KIND, **{k: v for k, v in c.items() if k in dict(_EncConfig._fields_)})
        self._h = c_void_p()
        _check(self.lib.foley_encoder_create(ctypes.byref(self.cfg), self.device.index or 0, ctypes.byref(self._h)))
        self._finalized = False

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.foley_encoder_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_tensor(self, name, t):
        t = t.detach()
        if t.dtype not in (torch.bfloat16, torch.float32, torch.float16):
            return                                   # integer buffers (position_ids, token_type_ids) are not weights
        t = t.contiguous()
        shape = (c_int64 * max(t.dim(), 1))(*t.shape)
        _check(self.lib.foley_encoder_load_tensor(self._h, name.encode(), c_void_p(t.data_ptr()), shape, t.dim(), FOLEY_DT[t.dtype]))

    def load_state_dict(self, sd, prefix=""):
        """sd: the HF module's state dict (SiglipModel / SiglipVisionModel: keys `vision_model.*`;
        ClapTextModelWithProjection: keys `text_model.*`).  `prefix` is prepended to every key (e.g. "vision_model." for
        a bare SiglipVisionTransformer)."""
        for k, v in sd.items():
            if isinstance(v, torch.Tensor):
                self.load_tensor(prefix + k, v)
        return self

    def load_safetensors(self, path, prefix=""):
        n = c_int64()
        _check(self.lib.foley_encoder_load_safetensors(self._h, os.fsencode(path), prefix.encode(), ctypes.byref(n)))
        return n.value

    def finalize(self):
        _check(self.lib.foley_encoder_finalize(self._h))
        self._finalized = True
        return self

    def set_option(self, key, value):
        _check(self.lib.foley_encoder_set_option(self._h, key.encode(), int(value)))

    def launch_count(self):
        return int(self.lib.foley_encoder_launch_count(self._h))

    def debug_read(self, what, numel):
        out = torch.empty(int(numel), dtype=torch.float32)
        n = c_int64()
        _check(self.lib.foley_encoder_debug_read(self._h, what.encode(), c_void_p(out.data_ptr()), out.numel(), ctypes.byref(n)))
        return out[: n.value]


class SiglipVisionEncoder(_Encoder):
    """HF SiglipVisionTransformer (+ attention-pooling head) of google/siglip2-base-patch16-512 on the engine's kernels."""
    KIND = ENC_SIGLIP_VISION
    DEFAULTS = SIGLIP2_BASE_512

    @classmethod
    def from_hf(cls, model, device=None):
        """model: transformers SiglipModel / SiglipVisionModel (any dtype; weights are rounded to bf16 like `.to(bf16)`)."""
        vc = getattr(model.config, "vision_config", model.config)
        cfg = dict(hidden_size=vc.hidden_size, num_heads=vc.num_attention_heads, num_layers=vc.num_hidden_layers,
                   intermediate_size=vc.intermediate_size, layer_norm_eps=vc.layer_norm_eps, image_size=vc.image_size,
                   patch_size=vc.patch_size)
        if getattr(vc, "hidden_act", "gelu_pytorch_tanh") != "gelu_pytorch_tanh":
            raise FoleyError(f"SigLIP hidden_act {vc.hidden_act!r} is not supported (gelu_pytorch_tanh only)")
        enc = cls(cfg, device=device)
        enc.load_state_dict(model.state_dict())
        return enc.finalize()

    def encode(self, pixels):
        """pixels: fp32 [T, 3, image, image] on this device (preprocess.preprocess_video's 8 fps output) -> bf16 [T, hidden]."""
        if not self._finalized:
            self.finalize()
        px = pixels.to(self.device, torch.float32).contiguous()
        s = self.cfg.image_size
        if px.dim() != 4 or tuple(px.shape[1:]) != (3, s, s):
            raise FoleyError(f"pixels have shape {tuple(px.shape)}, expected [T, 3, {s}, {s}]")
        out = torch.empty(px.shape[0], self.cfg.hidden_size, dtype=torch.bfloat16, device=self.device)
        _check(self.lib.foley_siglip_encode(self._h, c_void_p(px.data_ptr()), px.shape[0], c_void_p(out.data_ptr()),
                                            _stream_ptr(self.device)))
        return out


class ClapTextEncoder(_Encoder):
    """HF ClapTextModel (RoBERTa) of laion/larger_clap_general on the engine's kernels: last_hidden_state."""
    KIND = ENC_CLAP_TEXT
    DEFAULTS = CLAP_TEXT_GENERAL

    @classmethod
    def from_hf(cls, model, device=None):
        """model: transformers ClapTextModelWithProjection (or ClapTextModel with prefix-less keys)."""
        tc = model.config
        if getattr(tc, "hidden_act", "gelu") != "gelu":
            raise FoleyError(f"CLAP hidden_act {tc.hidden_act!r} is not supported (gelu only)")
        cfg = dict(hidden_size=tc.hidden_size, num_heads=tc.num_attention_heads, num_layers=tc.num_hidden_layers,
                   intermediate_size=tc.intermediate_size, layer_norm_eps=tc.layer_norm_eps, vocab_size=tc.vocab_size,
                   max_positions=tc.max_position_embeddings, pad_token_id=tc.pad_token_id)
        enc = cls(cfg, device=device)
        sd = model.state_dict()
        prefix = "" if any(k.startswith("text_model.") for k in sd) else "text_model."
        enc.load_state_dict(sd, prefix=prefix)
        return enc.finalize()

    def encode(self, input_ids, attention_mask=None):
        """input_ids / attention_mask: integer [B, T] (tokenizer output, any device) -> bf16 [B, T, hidden] on this device."""
        if not self._finalized:
            self.finalize()
        ids = input_ids.detach().to("cpu", torch.int32).contiguous()
        B, T = ids.shape
        mask_p = None
        if attention_mask is not None:
            mask = attention_mask.detach().to("cpu", torch.int32).contiguous()
            if tuple(mask.shape) != (B, T):
                raise FoleyError("attention_mask must have the shape of input_ids")
            mask_p = ctypes.cast(mask.data_ptr(), POINTER(c_int32))
        out = torch.empty(B, T, self.cfg.hidden_size, dtype=torch.bfloat16, device=self.device)
        _check(self.lib.foley_clap_text_encode(self._h, ctypes.cast(ids.data_ptr(), POINTER(c_int32)), mask_p, B, T,
                                               c_void_p(out.data_ptr()), _stream_ptr(self.device)))
        return out


class SynchformerEncoder(_Encoder):
    """The Synchformer visual extractor (reference models/synchformer: MotionFormer, divided space-time attention + spatial
    aggregation layer) on the engine's kernels, with the arithmetic of the reference's call: fp16 autocast around a module
    whose parameters are in the DiT's dtype (feature_utils.py:81-106, nodes.py:283-284)."""
    KIND = ENC_SYNCHFORMER
    DEFAULTS = MOTIONFORMER_DIVIDED_224

    @classmethod
    def from_state_dict(cls, sd, device=None, **cfg):
        """sd: Synchformer state dict (keys `vfeat_extractor.*`; the audio extractor / sync head are ignored) or a bare
        MotionFormer state dict."""
        enc = cls(cfg or None, device=device)
        enc.load_state_dict(sd)
        return enc.finalize()

    def encode(self, frames):
        """frames: fp32 [T25, 3, 224, 224] on this device (preprocess.preprocess_video's 25 fps output) -> fp32
        [1, segments * 8, hidden], segments = (T25 - 16) // 8 + 1 (encode_video_with_sync's `b (s t) d`)."""
        if not self._finalized:
            self.finalize()
        fr = frames.to(self.device, torch.float32).contiguous()
        if fr.dim() != 4 or tuple(fr.shape[1:]) != (3, 224, 224):
            raise FoleyError(f"frames have shape {tuple(fr.shape)}, expected [T, 3, 224, 224]")
        if fr.shape[0] < 16:
            raise FoleyError("the Synchformer needs at least 16 frames (one 0.64 s window)")
        S = (fr.shape[0] - 16) // 8 + 1
        out = torch.empty(1, S * 8, self.cfg.hidden_size, dtype=torch.float32, device=self.device)
        _check(self.lib.foley_synchformer_encode(self._h, c_void_p(fr.data_ptr()), fr.shape[0], c_void_p(out.data_ptr()),
                                                 _stream_ptr(self.device)))
        return out


def attention_d64(q, k, v, heads, key_mask=None, round_scores=False, impl=0, scale=0.125):
    """foley_attention_d64 on [B, S, heads*64]-shaped views (any row / batch strides, last dim contiguous): returns bf16
    [B, Sq, heads*64]."""
    lib = _lib()
    for t in (q, k, v):
        assert t.dtype == torch.bfloat16 and t.is_cuda and t.dim() == 3 and t.stride(2) == 1
    B, Sq, Sk = q.shape[0], q.shape[1], k.shape[1]
    out = torch.empty(B, Sq, heads * 64, dtype=torch.bfloat16, device=q.device)
    km = key_mask.to(q.device, torch.int32).contiguous() if key_mask is not None else None
    _check(lib.foley_attention_d64(c_void_p(q.data_ptr()), c_void_p(k.data_ptr()), c_void_p(v.data_ptr()), c_void_p(out.data_ptr()),
                                   B, heads, Sq, Sk, q.stride(0), q.stride(1), k.stride(0), k.stride(1), out.stride(0),
                                   out.stride(1), float(scale), c_void_p(km.data_ptr()) if km is not None else None,
                                   1 if round_scores else 0, impl, _stream_ptr(q.device)))
    return out
