"""Which types does the reference's Synchformer call produce (feature_utils.py:100-102: fp16 autocast around a module that the
Sampler has moved to the DiT's dtype, nodes.py:283-284)?  The MotionFormer concatenates its class token (module dtype) with
the patch embeddings (fp16 under autocast) — video_model_builder.py forward_features: torch.cat((cls_tokens, x), dim=1); `cat`
is an autocast "promote" op.  Measured on the B200: bf16 module -> conv fp16, concatenation fp32 (the residual stream is fp32
from there on); fp16 module -> fp16.  Prints one JSON line (profiles/r02_autocast_probe.json)."""
import json

import torch

out = {}
for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16), ("fp32", torch.float32)):
    cls = torch.zeros(1, 1, 8, device="cuda", dtype=dt)
    conv = torch.nn.Conv3d(3, 8, (2, 16, 16), (2, 16, 16)).cuda().to(dt)
    x = torch.zeros(1, 3, 2, 16, 16, device="cuda")
    try:
        with torch.autocast(device_type="cuda", enabled=True, dtype=torch.half):
            y = conv(x).flatten(2).transpose(1, 2)
            z = torch.cat((cls, y), dim=1)
        out[name] = f"ok: conv -> {y.dtype}, cat -> {z.dtype}"
    except Exception as e:   # noqa: BLE001
        out[name] = "raises: " + str(e).splitlines()[0]
print(json.dumps({"torch": torch.__version__, "module dtype -> fp16-autocast cls/patch concat": out}))
