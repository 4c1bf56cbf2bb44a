"""Records the torch CUDA semantics the reference inherits on this box (dtype flow under
torch.autocast(bf16), nn.RMSNorm(eps=None) behaviour, TF32 defaults for conv1d) — generic torch ops
only, the reference itself is not present on the GPU box.  Output: gpurun_out/torch_probe.json.
The oracle's "cuda_bf16" policy (oracle/foley_oracle.py) is written against these facts.
"""
import json
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

out = {}
dev = "cuda"
torch.manual_seed(0)
out["torch"] = torch.__version__
out["gpu"] = torch.cuda.get_device_name(0)
out["cudnn_allow_tf32"] = torch.backends.cudnn.allow_tf32
out["matmul_allow_tf32"] = torch.backends.cuda.matmul.allow_tf32

# --- dtype flow under autocast
with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
    x = torch.randn(2, 16, 64, device=dev).bfloat16()
    ln = nn.LayerNorm(64, elementwise_affine=False, eps=1e-6).to(dev)
    out["layer_norm_out"] = str(ln(x).dtype)
    y = F.interpolate(x.transpose(1, 2), size=50, mode="nearest-exact")
    out["interp_nearest_exact_out"] = str(y.dtype)
    out["bf16_plus_f32"] = str((x + x.float()).dtype)
    lin = nn.Linear(64, 64).to(dev).bfloat16()
    out["linear_of_f32_input"] = str(lin(x.float()).dtype)
    rms = nn.RMSNorm(128).to(dev).bfloat16()
    q = torch.randn(2, 4, 16, 128, device=dev).bfloat16()
    out["nn_rmsnorm_out_autocast"] = str(rms(q).dtype)
    out["silu_bf16"] = str(F.silu(x).dtype)
    out["silu_f32"] = str(F.silu(x.float()).dtype)
    qf = torch.randn(2, 4, 16, 128, device=dev)
    kf = torch.randn(2, 4, 16, 128, device=dev)
    vb = torch.randn(2, 4, 16, 128, device=dev).bfloat16()
    out["sdpa_mixed_out"] = str(F.scaled_dot_product_attention(qf, kf, vb).dtype)
    c1 = nn.Conv1d(64, 32, 3, padding=1).to(dev).bfloat16()
    out["conv1d_out"] = str(c1(x.transpose(1, 2)).dtype)
    cat = torch.cat((x, x.float()), dim=1)
    out["cat_bf16_f32"] = str(cat.dtype)

# --- nn.RMSNorm(eps=None) on bf16: which eps, which rounding
q = (torch.randn(64, 12, 250, 128, device=dev) * 0.3).bfloat16()
w = (1.0 + 0.1 * torch.randn(128, device=dev)).bfloat16()
rms = nn.RMSNorm(128).to(dev).bfloat16()
with torch.no_grad():
    rms.weight.copy_(w)
got = rms(q)
qf = q.float()
ms = qf.pow(2).mean(-1, keepdim=True)
cands = {}
for ename, eps in [("eps_bf16", torch.finfo(torch.bfloat16).eps), ("eps_f32", torch.finfo(torch.float32).eps),
                   ("eps_1e-6", 1e-6), ("eps_0", 0.0)]:
    n = qf * torch.rsqrt(ms + eps)
    cands[ename + "/round_then_wmul"] = n.bfloat16() * w
    cands[ename + "/single_round"] = (n * w.float()).bfloat16()
res = {}
for k, v in cands.items():
    res[k] = {"mismatch_frac": (v != got).float().mean().item(),
              "rel": ((v.float() - got.float()).norm() / got.float().norm()).item()}
out["nn_rmsnorm_bf16"] = res
out["nn_rmsnorm_bf16_best"] = min(res, key=lambda k: res[k]["mismatch_frac"])

# the reference's custom RMSNorm (norm_layers.py:49-51) for comparison of formulas only
n6 = (qf * torch.rsqrt(ms + 1e-6)).bfloat16() * w
out["custom_rms_vs_nn_rms_mismatch"] = (n6 != got).float().mean().item()

# --- TF32 in fp32 conv1d (DAC path: nodes.py:398 moves DAC to fp32; cudnn.allow_tf32 default)
xc = torch.randn(1, 512, 4000, device=dev)
conv = nn.Conv1d(512, 512, 7, padding=3).to(dev)
ref64 = F.conv1d(xc.double(), conv.weight.double(), conv.bias.double(), padding=3)
y_def = conv(xc)
out["conv1d_f32_default_rel_vs_f64"] = ((y_def.double() - ref64).norm() / ref64.norm()).item()
torch.backends.cudnn.allow_tf32 = False
y_no = conv(xc)
out["conv1d_f32_tf32off_rel_vs_f64"] = ((y_no.double() - ref64).norm() / ref64.norm()).item()
torch.backends.cudnn.allow_tf32 = True
ct = nn.ConvTranspose1d(512, 256, 8, stride=4, padding=2).to(dev)
ref64 = F.conv_transpose1d(xc.double(), ct.weight.double(), ct.bias.double(), stride=4, padding=2)
out["convT_f32_default_rel_vs_f64"] = ((ct(xc).double() - ref64).norm() / ref64.norm()).item()

# --- which SDPA backend handles bf16 [2,12,290,128]
qb = torch.randn(2, 12, 290, 128, device=dev).bfloat16()
ref = F.scaled_dot_product_attention(qb.float(), qb.float(), qb.float())
got = F.scaled_dot_product_attention(qb, qb, qb)
out["sdpa_bf16_rel_vs_f32"] = ((got.float() - ref).norm() / ref.norm()).item()

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/torch_probe.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
