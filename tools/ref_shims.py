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


def install_synchformer_shims():
    """omegaconf + timm stand-ins, enough for the reference's MotionFormer (models/synchformer/motionformer.py,
    video_model_builder.py, vit_helper.py): OmegaConf.load of the committed divided_224_16x4.yaml into an attribute dict that
    accepts assignment, timm.layers.{trunc_normal_, to_2tuple}."""
    if "omegaconf" not in sys.modules:
        import yaml

        class _Node(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

            def __setattr__(self, k, v):
                self[k] = v

        def _wrap(o):
            return _Node({k: _wrap(v) for k, v in o.items()}) if isinstance(o, dict) else o

        oc = _mod("omegaconf")

        class OmegaConf:
            overrides = {}          # e.g. {"VIT.DEPTH": 2}: applied to every loaded config (small-depth goldens)

            @staticmethod
            def load(path):
                with open(path) as f:
                    cfg = _wrap(yaml.safe_load(f))
                for k, v in OmegaConf.overrides.items():
                    node = cfg
                    parts = k.split(".")
                    for q in parts[:-1]:
                        node = node[q]
                    node[parts[-1]] = v
                return cfg

        oc.OmegaConf = OmegaConf
    try:
        import timm  # noqa: F401
    except Exception:   # noqa: BLE001
        timm = _mod("timm")
        layers = _mod("timm.layers")
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        timm.layers = layers
        tm = _mod("timm.models")
        tml = _mod("timm.models.layers")
        tml.trunc_normal_, tml.to_2tuple = layers.trunc_normal_, layers.to_2tuple
        timm.models, tm.layers = tm, tml


def load_motionformer(depth=None):
    """The reference's MotionFormer class exactly as Synchformer.__init__ builds it (synchformer.py:21-27); `depth` overrides
    VIT.DEPTH of its YAML (small-depth goldens)."""
    install_synchformer_shims()
    sys.modules["omegaconf"].OmegaConf.overrides = {"VIT.DEPTH": depth} if depth else {}
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # import the sub-module without the package __init__ chain (utils/feature_utils pull av, loguru, ...)
    pkg = "hunyuanvideo_foley.models.synchformer"
    for name in ("hunyuanvideo_foley", "hunyuanvideo_foley.models", pkg):
        if name not in sys.modules:
            m = _mod(name)
            m.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
    mf = importlib.import_module(pkg + ".motionformer")
    return lambda: mf.MotionFormer(extract_features=True, factorize_space_time=True, agg_space_module="TransformerEncoderLayer",
                                   agg_time_module="torch.nn.Identity", add_global_repr=False)
