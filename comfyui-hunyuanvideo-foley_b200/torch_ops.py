"""`torch.ops.foley_b200.*` — the C-ABI entry points of libfoley_b200.so surfaced as torch ops (SURVEY.md §8b,
north-star: "through a thin C-ABI extension surfaced as a torch op"), so the Sampler body stays a few lines:

    torch.ops.foley_b200.set_conditions(h, clip, sync, text, L, batch)     foley_set_conditions
    torch.ops.foley_b200.dit_forward(h, x, t)            -> Tensor          foley_dit_forward   (hifi_foley.py:707-924)
    torch.ops.foley_b200.denoise(h, latents, sigmas, g)  -> Tensor          foley_denoise       (utils.py:203-247)
    torch.ops.foley_b200.denoise_solver(h, latents, sigmas, g, solver) -> Tensor   foley_denoise_solver
    torch.ops.foley_b200.dac_decode(h, z)                -> Tensor          foley_dac_decode    (dac.py:280-303)

`h` is an integer handle from `register_engine(FoleyEngine)`.  Only the CUDA dispatch key has a kernel: calling an
op with CPU tensors raises (there is no CPU / eager path), the Meta key gives shapes for tracing tools.
"""
import threading
import weakref

import torch

from .engine import FoleyEngine, FoleyError

NAMESPACE = "foley_b200"
SOLVER_IDS = {"euler": 0, "heun-2": 1, "midpoint-2": 2, "kutta-4": 3}

_engines = weakref.WeakValueDictionary()
_next_handle = [1]
_handle_lock = threading.Lock()   # the multi-GPU Sampler registers one engine per host thread (parallel.denoise_sharded)


def register_engine(engine):
    """Returns the integer handle the ops take for `engine` (idempotent, thread-safe)."""
    with _handle_lock:
        h = getattr(engine, "_op_handle", None)
        if h is None or _engines.get(h) is not engine:
            h = _next_handle[0]
            _next_handle[0] += 1
            engine._op_handle = h
            _engines[h] = engine
        return h


def _engine(handle):
    eng = _engines.get(int(handle))
    if eng is None:
        raise FoleyError(f"foley_b200: unknown engine handle {handle}")
    return eng


_lib = torch.library.Library(NAMESPACE, "DEF")
_lib.define("set_conditions(int engine, Tensor clip, Tensor sync, Tensor text, int L, int batch) -> ()")
_lib.define("dit_forward(int engine, Tensor x, Tensor t) -> Tensor")
_lib.define("denoise(int engine, Tensor latents, Tensor sigmas, float guidance) -> Tensor")
_lib.define("denoise_solver(int engine, Tensor latents, Tensor sigmas, float guidance, int solver) -> Tensor")
_lib.define("dac_decode(int engine, Tensor z) -> Tensor")


# ---- CUDA kernels: thin calls into the ctypes binding (engine.py) ----------------------------------------------
def _set_conditions_cuda(engine, clip, sync, text, L, batch):
    _engine(engine).set_conditions(clip, sync, text, L=L, batch=batch)


def _dit_forward_cuda(engine, x, t):
    return _engine(engine).dit_forward(x, t)


def _denoise_cuda(engine, latents, sigmas, guidance):
    eng = _engine(engine)
    return eng.denoise(latents, sigmas, guidance, progress=getattr(eng, "_progress", None))


def _denoise_solver_cuda(engine, latents, sigmas, guidance, solver):
    eng = _engine(engine)
    return eng.denoise_solver(latents, sigmas, guidance, solver, progress=getattr(eng, "_progress", None))


def _dac_decode_cuda(engine, z):
    return _engine(engine).dac_decode(z)


_lib.impl("set_conditions", _set_conditions_cuda, "CUDA")
_lib.impl("dit_forward", _dit_forward_cuda, "CUDA")
_lib.impl("denoise", _denoise_cuda, "CUDA")
_lib.impl("denoise_solver", _denoise_solver_cuda, "CUDA")
_lib.impl("dac_decode", _dac_decode_cuda, "CUDA")


# ---- CPU key: fail loudly (the product path is the CUDA library; no fallback) ----------------------------------
def _no_cpu(*_args, **_kw):
    raise FoleyError("foley_b200 ops have no CPU implementation: tensors must live on an sm_100a CUDA device")


for _name in ("set_conditions", "dit_forward", "denoise", "denoise_solver", "dac_decode"):
    _lib.impl(_name, _no_cpu, "CPU")


# ---- Meta key: output shapes only -----------------------------------------------------------------------------
def _same_f32_meta(t):
    return torch.empty(t.shape, dtype=torch.float32, device="meta")


_lib.impl("set_conditions", lambda engine, clip, sync, text, L, batch: None, "Meta")
_lib.impl("dit_forward", lambda engine, x, t: _same_f32_meta(x), "Meta")
_lib.impl("denoise", lambda engine, latents, sigmas, guidance: _same_f32_meta(latents), "Meta")
_lib.impl("denoise_solver", lambda engine, latents, sigmas, guidance, solver: _same_f32_meta(latents), "Meta")
_lib.impl("dac_decode",
          lambda engine, z: torch.empty(z.shape[0], 1, z.shape[2] * 960, dtype=torch.float32, device="meta"), "Meta")


# ---- convenience wrappers used by sampling.py -----------------------------------------------------------------
def set_conditions(engine: FoleyEngine, clip, sync, text, L, batch):
    torch.ops.foley_b200.set_conditions(register_engine(engine), clip, sync, text, int(L), int(batch))


def dit_forward(engine: FoleyEngine, x, t):
    t = torch.as_tensor(t, dtype=torch.float32).flatten()
    return torch.ops.foley_b200.dit_forward(register_engine(engine), x, t)


def denoise(engine: FoleyEngine, latents, sigmas, guidance, solver="euler", progress=None):
    """Whole sampling loop in the engine.  `progress(step)` (ComfyUI ProgressBar, utils.py:247) cannot travel through
    an op schema, so it is parked on the engine object for the duration of the call."""
    if solver not in SOLVER_IDS:
        raise ValueError(f"Solver {solver} not supported. Supported solvers: {list(SOLVER_IDS)}")
    h = register_engine(engine)
    sig = torch.as_tensor(sigmas, dtype=torch.float32).flatten()
    engine._progress = progress
    try:
        if solver == "euler":
            return torch.ops.foley_b200.denoise(h, latents, sig, float(guidance))
        return torch.ops.foley_b200.denoise_solver(h, latents, sig, float(guidance), SOLVER_IDS[solver])
    finally:
        engine._progress = None


def dac_decode(engine: FoleyEngine, z):
    return torch.ops.foley_b200.dac_decode(register_engine(engine), z)
