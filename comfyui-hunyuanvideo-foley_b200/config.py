"""YAML config loading with attribute access — the interface the reference's nodes use
(`cfg.model_config.model_kwargs.audio_frame_rate`, `cfg.diffusion_config.sample_flow_shift`;
reference hunyuanvideo_foley/utils/config_utils.py:7-109, nodes.py:79-82,269-271)."""
import os

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
CONFIG_DIR = os.path.join(_HERE, "configs")


class AttributeDict(dict):
    """dict whose keys are also attributes, recursively; `.get` works as for dicts."""

    def __init__(self, data=None):
        super().__init__()
        for k, v in (data or {}).items():
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, AttributeDict):
            return cls(v)
        if isinstance(v, list):
            return [cls._wrap(x) for x in v]
        return v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = self._wrap(v)


def load_yaml(path):
    if not os.path.exists(path):
        raise FileNotFoundError(f"Hunyuan config file not found at {path}")
    with open(path, "r", encoding="utf-8") as f:
        return AttributeDict(yaml.safe_load(f))


def config_path(model_size="xxl"):
    return os.path.join(CONFIG_DIR, f"hunyuanvideo-foley-{model_size}.yaml")


def load_model_config(model_size="xxl"):
    """The Sampler always loads the XXL YAML (reference nodes.py:79,269); XL is reachable from the library."""
    return load_yaml(config_path(model_size))


def detect_model_size(state_dict):
    """Picks xl / xxl from the checkpoint's hidden size (the reference never does: its loader hard-codes XXL)."""
    w = state_dict.get("audio_embedder.proj.weight")
    if w is not None and int(w.shape[0]) == 1408:
        return "xl"
    return "xxl"
