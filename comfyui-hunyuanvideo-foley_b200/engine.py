"""Python binding of libfoley_b200.so (include/foley_b200.h) — ctypes over the C ABI, torch only for
device memory and streams.  The CUDA library is the product path: if it is missing or fails to load this
module raises, there is no eager / CPU fallback.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FOLEY_B200_LIB", os.path.join(_HERE, "libfoley_b200.so"))   # override: A/B builds only

FOLEY_DT = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2, torch.float8_e4m3fn: 3, torch.float8_e5m2: 4}
FP8_STORAGE = {"none": 0, None: 0, "fp8_e4m3fn": 1, "fp8_e5m2": 2}   # option "fp8_weight_storage"


class FoleyError(RuntimeError):
    """Raised for any non-zero foley_status (the reference raises RuntimeError on bad shapes)."""


class _Config(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("hidden_size", "num_heads", "depth_triple_blocks", "depth_single_blocks",
                                        "mlp_hidden_triple", "mlp_hidden_single", "sync_hidden", "latent_dim",
                                        "clip_dim", "sync_dim", "text_dim", "freq_dim")] + \
               [("rope_theta", c_float), ("single_rms_eps", c_float)] + \
               [(n, c_int32) for n in ("max_batch", "max_seconds", "with_dac")]


class _AttnSrc(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("batch_stride", c_int64), ("head_stride", c_int64), ("row_stride", c_int64),
                ("rows", c_int32), ("rows0", c_int32), ("batch", c_int32), ("reserved", c_int32),
                ("norm_w", c_void_p * 2), ("rope", c_void_p * 2)]


class _AttnArgs(ctypes.Structure):
    _fields_ = [("q", _AttnSrc), ("k", _AttnSrc), ("v", _AttnSrc), ("out", c_void_p), ("out_batch_stride", c_int64),
                ("batch", c_int32), ("heads", c_int32), ("kv_batch_map", c_void_p),
                ("scale", c_float), ("norm_kind", c_int32), ("eps", c_float), ("impl", c_int32), ("dbg", c_int32 * 4)]


PROGRESS_FN = ctypes.CFUNCTYPE(None, c_int32, c_void_p)
_lib = None


def load_library(path=None):
    """dlopen the engine; raises FoleyError with the build hint when the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FoleyError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         f"(make -C {os.path.join(_HERE, 'csrc')}); there is no fallback path")
    lib = ctypes.CDLL(path)
    lib.foley_last_error.restype = c_char_p
    lib.foley_version.restype = c_char_p
    lib.foley_engine_create.argtypes = [POINTER(_Config), c_int32, POINTER(c_void_p)]
    lib.foley_engine_destroy.argtypes = [c_void_p]
    lib.foley_engine_destroy.restype = None
    lib.foley_engine_load_tensor.argtypes = [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int32, c_int32]
    lib.foley_engine_load_safetensors.argtypes = [c_void_p, c_char_p, c_char_p, POINTER(c_int64)]
    lib.foley_safetensors_probe.argtypes = [c_char_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]
    lib.foley_fp8_wraps.argtypes = [c_char_p, c_int32]
    lib.foley_fp8_wraps.restype = c_int32
    lib.foley_preprocess_frames.argtypes = [c_void_p, c_int32, c_int32, c_int32, POINTER(c_int32), c_int32, c_int32, c_int32,
                                            c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.foley_resize_weights.argtypes = [c_int32, c_int32, POINTER(c_int32), POINTER(c_int32), POINTER(ctypes.c_int16),
                                         c_int64, POINTER(c_int32), POINTER(c_int32)]
    lib.foley_engine_finalize.argtypes = [c_void_p]
    lib.foley_set_conditions.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                         c_int32, c_int32, c_int32, c_void_p]
    lib.foley_dit_forward.argtypes = [c_void_p, c_void_p, POINTER(c_float), c_int32, c_void_p, c_void_p]
    lib.foley_denoise.argtypes = [c_void_p, c_void_p, POINTER(c_float), c_int32, c_float, PROGRESS_FN, c_void_p,
                                  c_void_p]
    lib.foley_denoise_solver.argtypes = [c_void_p, c_void_p, POINTER(c_float), c_int32, c_float, c_int32, PROGRESS_FN,
                                         c_void_p, c_void_p]
    lib.foley_dac_decode.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]
    lib.foley_launch_count.argtypes = [c_void_p]
    lib.foley_launch_count.restype = c_int64
    lib.foley_debug_read.argtypes = [c_void_p, c_char_p, c_void_p, c_int64, POINTER(c_int64)]
    lib.foley_debug_flags.argtypes = [POINTER(ctypes.c_uint32)]
    lib.foley_engine_set_option.argtypes = [c_void_p, c_char_p, c_int64]
    lib.foley_attention.argtypes = [POINTER(_AttnArgs), c_void_p]
    _lib = lib
    return lib


def _check(status):
    if status != 0:
        raise FoleyError(f"foley_b200 error {status}: {_lib.foley_last_error().decode()}")


def _stream_ptr(device):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def engine_config(model_cfg, with_dac=True, max_batch=8, max_seconds=60):
    """model_cfg: dict with the YAML's model_kwargs keys (see config.py)."""
    C = int(model_cfg["hidden_size"])

    def convmlp_hidden(h, multiple=256):          # mlp_layers.py:133-134
        h = int(2 * h / 3)
        return multiple * ((h + multiple - 1) // multiple)

    mlp_ratio = model_cfg.get("mlp_ratio", 4)
    return _Config(hidden_size=C, num_heads=int(model_cfg["num_heads"]),
                   depth_triple_blocks=int(model_cfg["depth_triple_blocks"]),
                   depth_single_blocks=int(model_cfg["depth_single_blocks"]),
                   mlp_hidden_triple=int(C * mlp_ratio), mlp_hidden_single=convmlp_hidden(C * mlp_ratio),
                   sync_hidden=convmlp_hidden(C * 4), latent_dim=int(model_cfg.get("audio_vae_latent_dim", 128)),
                   clip_dim=int(model_cfg.get("clip_dim", 768)), sync_dim=int(model_cfg.get("sync_feat_dim", 768)),
                   text_dim=int(model_cfg.get("condition_dim", 768)), freq_dim=256,
                   rope_theta=float(model_cfg.get("rope_theta", 10000)),
                   single_rms_eps=float(torch.finfo(torch.float32).eps),   # profiles/r01_torch_probe.json
                   max_batch=max_batch, max_seconds=max_seconds, with_dac=1 if with_dac else 0)


class FoleyEngine:
    """One engine instance per GPU (one process per GPU)."""

    def __init__(self, model_cfg, device=None, with_dac=False):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise FoleyError("foley_b200 needs a CUDA device (sm_100a); no CPU path exists")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.model_cfg = dict(model_cfg)
        self.cfg = engine_config(model_cfg, with_dac=with_dac)
        self._h = c_void_p()
        _check(self.lib.foley_engine_create(ctypes.byref(self.cfg), self.device.index or 0, ctypes.byref(self._h)))
        self.latent_dim = self.cfg.latent_dim
        self.plan = None
        self._finalized = False

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.foley_engine_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights
    def set_fp8_weight_storage(self, quantization):
        """The reference's `quantization` choice ("none" / "fp8_e4m3fn" / "fp8_e5m2", nodes.py:66,106-121) for the
        tensors loaded next: the weights it would keep in FP8 are rounded through that format on the device."""
        if quantization not in FP8_STORAGE:
            raise ValueError(f"unknown quantization {quantization!r}")
        self.set_option("fp8_weight_storage", FP8_STORAGE[quantization])

    def load_safetensors(self, path, prefix=""):
        """The whole .safetensors checkpoint, file -> device, no host state dict (foley_engine_load_safetensors)."""
        n = c_int64()
        _check(self.lib.foley_engine_load_safetensors(self._h, os.fsencode(path), prefix.encode(), ctypes.byref(n)))
        self._finalized = False
        return n.value

    def load_tensor(self, name, t):
        t = t.detach()
        if t.dtype not in FOLEY_DT:
            t = t.float()
        t = t.contiguous()   # fp8 checkpoint tensors travel as their bytes and are de-quantised on the device
        shape = (c_int64 * max(t.dim(), 1))(*t.shape)
        _check(self.lib.foley_engine_load_tensor(self._h, name.encode(), c_void_p(t.data_ptr()), shape, t.dim(),
                                                 FOLEY_DT[t.dtype]))

    def load_state_dict(self, sd, prefix=""):
        """sd: reference state dict (HunyuanVideoFoley.state_dict(), or DAC's with prefix='dac.')."""
        for k, v in sd.items():
            if isinstance(v, torch.Tensor):
                self.load_tensor(prefix + k, v)
        self._finalized = False

    def finalize(self):
        _check(self.lib.foley_engine_finalize(self._h))
        self._finalized = True

    # ---- conditions
    def set_conditions(self, clip, sync, text, L, batch):
        """clip [U,Lv,768], sync [U,S,768], text [U,T,768] on this device; U = 2 (uncond first) or 1."""
        if not self._finalized:
            self.finalize()
        U = clip.shape[0]
        dt = clip.dtype if clip.dtype in (torch.bfloat16, torch.float32, torch.float16) else torch.float32
        clip, sync, text = (x.to(self.device, dt).contiguous() for x in (clip, sync, text))
        if sync.shape[0] != U or text.shape[0] != U:
            raise FoleyError("clip / sync / text must share the leading n_cond dimension")
        _check(self.lib.foley_set_conditions(self._h, c_void_p(clip.data_ptr()), c_void_p(sync.data_ptr()),
                                             c_void_p(text.data_ptr()), FOLEY_DT[dt], U, clip.shape[1],
                                             sync.shape[1], text.shape[1], int(L), int(batch),
                                             _stream_ptr(self.device)))
        self.plan = dict(U=U, B=int(batch), L=int(L), Lv=clip.shape[1], S=sync.shape[1], T=text.shape[1])

    # ---- ops
    def dit_forward(self, x, t):
        """x [B*U, latent, L] (uncond rows first), t: float or [B*U] -> model output fp32 [B*U, latent, L]."""
        p = self.plan
        if p is None:
            raise FoleyError("set_conditions must be called before dit_forward")
        x = x.to(self.device, torch.float32).contiguous()
        if tuple(x.shape) != (p["B"] * p["U"], self.latent_dim, p["L"]):
            raise FoleyError(f"x has shape {tuple(x.shape)}, expected {(p['B'] * p['U'], self.latent_dim, p['L'])}")
        tt = torch.as_tensor(t, dtype=torch.float32).flatten().cpu()
        arr = (c_float * tt.numel())(*tt.tolist())
        out = torch.empty_like(x)
        _check(self.lib.foley_dit_forward(self._h, c_void_p(x.data_ptr()), arr, tt.numel(), c_void_p(out.data_ptr()),
                                          _stream_ptr(self.device)))
        return out

    SOLVERS = {"euler": 0, "heun-2": 1, "midpoint-2": 2, "kutta-4": 3}    # FOLEY_SOLVER_*

    def denoise_solver(self, latents, sigmas, guidance, solver, progress=None):
        """The reference loop with any of its four solvers (name or FOLEY_SOLVER_* id), whole loop in the engine."""
        sid = self.SOLVERS.get(solver, solver)
        if sid not in (0, 1, 2, 3):
            raise ValueError(f"Solver {solver} not supported. Supported solvers: {list(self.SOLVERS)}")
        return self.denoise(latents, sigmas, guidance, progress=progress, solver_id=int(sid))

    def denoise(self, latents, sigmas, guidance, progress=None, solver_id=0):
        """Euler loop in place on a copy of `latents` [B, latent, L]; returns fp32 latents on device."""
        p = self.plan
        if p is None:
            raise FoleyError("set_conditions must be called before denoise")
        lat = latents.to(self.device, torch.float32).contiguous().clone()
        if tuple(lat.shape) != (p["B"], self.latent_dim, p["L"]):
            raise FoleyError(f"latents have shape {tuple(lat.shape)}, expected {(p['B'], self.latent_dim, p['L'])}")
        sig = torch.as_tensor(sigmas, dtype=torch.float32).flatten().cpu()
        arr = (c_float * sig.numel())(*sig.tolist())
        cb = PROGRESS_FN(lambda step, _u: progress(step)) if progress is not None else PROGRESS_FN()
        if solver_id == 0:
            _check(self.lib.foley_denoise(self._h, c_void_p(lat.data_ptr()), arr, sig.numel() - 1, float(guidance), cb,
                                          None, _stream_ptr(self.device)))
        else:
            _check(self.lib.foley_denoise_solver(self._h, c_void_p(lat.data_ptr()), arr, sig.numel() - 1,
                                                 float(guidance), solver_id, cb, None, _stream_ptr(self.device)))
        return lat

    def dac_decode(self, z):
        """z [B, latent, L] -> waveform fp32 [B, 1, L*960] (dac.py:280-303)."""
        z = z.to(self.device, torch.float32).contiguous()
        B, _, L = z.shape
        wav = torch.empty(B, 1, L * 960, dtype=torch.float32, device=self.device)
        _check(self.lib.foley_dac_decode(self._h, c_void_p(z.data_ptr()), B, L, c_void_p(wav.data_ptr()),
                                         _stream_ptr(self.device)))
        return wav

    # ---- introspection
    def launch_count(self):
        return int(self.lib.foley_launch_count(self._h))

    def set_option(self, key, value):
        _check(self.lib.foley_engine_set_option(self._h, key.encode(), int(value)))

    def debug_read(self, what):
        n = c_int64()
        _check(self.lib.foley_debug_read(self._h, what.encode(), None, 0, ctypes.byref(n)))
        out = torch.empty(n.value, dtype=torch.float32, device=self.device)
        _check(self.lib.foley_debug_read(self._h, what.encode(), c_void_p(out.data_ptr()), n.value, ctypes.byref(n)))
        return out

    def debug_flags(self):
        buf = (ctypes.c_uint32 * 4)()
        _check(self.lib.foley_debug_flags(buf))
        return list(buf)


def attention(q, k, v, out, heads, kv_batch_map=None, scale=None, norm_kind=0, eps=1e-6, impl=0, dbg=(0, 0, 0, 0), stream=None):
    """foley_attention (include/foley_b200.h): softmax(Q K^T * scale) V for head_dim 128 on the tcgen05 kernel, with the
    q/k RMSNorm + RoPE optionally folded into the operand load.  q / k / v: operand dicts {t: bf16 CUDA tensor, off: element
    offset, batch_stride, head_stride, row_stride, rows, batch, [rows0, norm: [(w [128] bf16, rope [n,64,2] fp32) | None] * 2]};
    out: bf16 [batch, Sq, heads*128]."""
    lib = load_library()
    a = _AttnArgs()
    for name, op in (("q", q), ("k", k), ("v", v)):
        src = getattr(a, name)
        t = op["t"]
        assert t.dtype == torch.bfloat16 and t.is_cuda
        src.ptr = t.data_ptr() + 2 * op.get("off", 0)
        src.batch_stride, src.head_stride, src.row_stride = op["batch_stride"], op["head_stride"], op["row_stride"]
        src.rows, src.rows0, src.batch = op["rows"], op.get("rows0", op["rows"]), op["batch"]
        for i, nm in enumerate(op.get("norm", ())):
            if nm is not None:
                w, rope = nm
                assert rope.dtype == torch.float32 and rope.shape[-2:] == (64, 2) and rope.is_contiguous() and w.dtype == torch.bfloat16
                src.norm_w[i], src.rope[i] = w.data_ptr(), rope.data_ptr()
    a.out, a.out_batch_stride = out.data_ptr(), out.stride(0)
    a.batch, a.heads = out.shape[0], heads
    a.kv_batch_map = kv_batch_map.data_ptr() if kv_batch_map is not None else None
    a.scale = scale if scale is not None else 128 ** -0.5
    a.norm_kind, a.eps, a.impl = norm_kind, eps, impl
    for i in range(4):
        a.dbg[i] = dbg[i]
    _check(lib.foley_attention(ctypes.byref(a), stream))
    return out
