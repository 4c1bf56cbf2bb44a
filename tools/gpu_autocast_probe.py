"""Does the reference's Synchformer call (feature_utils.py:100-102: fp16 autocast around a module that the Sampler has
moved to the DiT's dtype, nodes.py:283-284) run at all when that dtype is bf16?  The MotionFormer concatenates its bf16 class
token with the fp16 patch embeddings (video_model_builder.py forward_features: torch.cat((cls_tokens, x), dim=1)) under
autocast; `cat` is an autocast "promote" op.  Prints one JSON line."""
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
