"""foley_b200 — B200-native HunyuanVideo-Foley denoising engine behind the ComfyUI node surface of
phazei/ComfyUI-HunyuanVideo-Foley (reference __init__.py:1-12: the package directory is put on sys.path and
the node mappings are re-exported so ComfyUI discovers the same six nodes)."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.append(_HERE)

from .nodes import NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS  # noqa: E402

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]
