"""Import shims that let the reference's hot path (/root/reference) be imported in THIS container,
where ComfyUI, diffusers, av, accelerate are not installed.

Only used by tools/make_golden.py / tools/gpu_reference_golden.py (fixture generation), bench.py --impl reference
and tests that are skipped when no reference tree is present.  Nothing under the product package imports this file.
The tree is looked up at $FOLEY_REFERENCE_ROOT, then baseline/_ref (tools/stage_reference.py), then /root/reference.
"""
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")   # tools/stage_reference.py (travels to the GPU box)


def _default_root():
    if os.environ.get("FOLEY_REFERENCE_ROOT"):
        return os.environ["FOLEY_REFERENCE_ROOT"]
    if os.path.isdir(os.path.join(_STAGED, "hunyuanvideo_foley")):
        return _STAGED
    return "/root/reference"


REF_ROOT = _default_root()


def _mod(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []
    sys.modules[name] = m
    return m


def install():
    if "comfy" in sys.modules and getattr(sys.modules["comfy"], "_foley_shim", False):
        return
    comfy = _mod("comfy")
    comfy._foley_shim = True
    mm = _mod("comfy.model_management")
    mm.get_torch_device = lambda: torch.device("cpu")
    mm.unet_offload_device = lambda: torch.device("cpu")
    mm.soft_empty_cache = lambda: None
    cu = _mod("comfy.utils")

    class ProgressBar:
        def __init__(self, total):
            self.total, self.n = total, 0

        def update(self, k):
            self.n += k

    cu.ProgressBar = ProgressBar
    cu.load_torch_file = lambda path, device=None: torch.load(path, map_location="cpu")
    comfy.model_management, comfy.utils = mm, cu
    fp = _mod("folder_paths")
    fp.models_dir = "/tmp/models"
    fp.folder_names_and_paths = {}
    fp.supported_pt_extensions = {".pt", ".pth", ".safetensors"}

    diffusers = _mod("diffusers")
    dmodels = _mod("diffusers.models")

    class ModelMixin(torch.nn.Module):
        @property
        def dtype(self):
            return next(self.parameters()).dtype

        @property
        def device(self):
            return next(self.parameters()).device

    dmodels.ModelMixin = ModelMixin
    dcfg = _mod("diffusers.configuration_utils")

    class ConfigMixin:
        config_name = "config.json"

    def register_to_config(fn):
        import functools
        import inspect

        @functools.wraps(fn)
        def wrapper(self, *a, **k):
            sig = inspect.signature(fn)
            bound = sig.bind(self, *a, **k)
            bound.apply_defaults()
            cfg = {n: v for n, v in bound.arguments.items() if n != "self"}
            self.config = types.SimpleNamespace(**cfg)
            return fn(self, *a, **k)

        return wrapper

    dcfg.ConfigMixin, dcfg.register_to_config = ConfigMixin, register_to_config
    dutils = _mod("diffusers.utils")

    class BaseOutput:
        """diffusers.utils.BaseOutput stand-in: dataclass subclasses index like tuples (utils.py:246)."""

        def to_tuple(self):
            import dataclasses
            return tuple(getattr(self, f.name) for f in dataclasses.fields(self))

        def __getitem__(self, k):
            return self.to_tuple()[k] if isinstance(k, int) else getattr(self, k)

    dutils.BaseOutput = BaseOutput

    class _Log:
        @staticmethod
        def get_logger(name):
            import logging
            return logging.getLogger(name)

    dutils.logging = _Log
    dtu = _mod("diffusers.utils.torch_utils")

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        # diffusers draws on the generator's device (CPU here) in the target dtype, then moves.
        gdev = generator.device if generator is not None else (device or torch.device("cpu"))
        t = torch.randn(shape, generator=generator, device=gdev, dtype=dtype)
        return t.to(device) if device is not None else t

    dtu.randn_tensor = randn_tensor
    dsch = _mod("diffusers.schedulers")
    dsch.DDPMScheduler = type("DDPMScheduler", (), {})
    dsch.EulerDiscreteScheduler = type("EulerDiscreteScheduler", (), {})
    dsu = _mod("diffusers.schedulers.scheduling_utils")
    dsu.SchedulerMixin = type("SchedulerMixin", (), {})
    diffusers.models, diffusers.utils, diffusers.schedulers = dmodels, dutils, dsch
    _mod("av")
    if "loguru" not in sys.modules:
        try:
            import loguru  # noqa: F401
        except Exception:
            lg = _mod("loguru")
            import logging
            lg.logger = logging.getLogger("foley-ref")

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # first import raises inside dac_vae (audiotools missing) and is swallowed; retry resolves
    for _ in range(2):
        try:
            importlib.import_module("hunyuanvideo_foley")
        except Exception:
            pass


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "hunyuanvideo_foley"))


def load_reference():
    """Returns a namespace with the reference classes/functions on the hot path."""
    install()
    ns = types.SimpleNamespace()
    from hunyuanvideo_foley.models.hifi_foley import HunyuanVideoFoley
    from hunyuanvideo_foley.utils.config_utils import load_yaml, AttributeDict
    from hunyuanvideo_foley.utils.schedulers import FlowMatchDiscreteScheduler
    try:
        from hunyuanvideo_foley.models.dac_vae.model.dac import DAC
    except Exception:
        from hunyuanvideo_foley.models.dac_vae.model.dac import DAC
    spec = importlib.util.spec_from_file_location("foley_ref_utils", os.path.join(REF_ROOT, "utils.py"))
    ref_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_utils)
    ns.HunyuanVideoFoley, ns.load_yaml, ns.AttributeDict = HunyuanVideoFoley, load_yaml, AttributeDict
    ns.FlowMatchDiscreteScheduler, ns.DAC, ns.utils = FlowMatchDiscreteScheduler, DAC, ref_utils
    ns.config_dir = os.path.join(REF_ROOT, "configs")
    return ns
