"""TEST INFRASTRUCTURE — CPU oracle for the condition encoders.  NOT a product path: only tests/,
__graft_entry__.smoke() and bench.py may import it.

Functional fp32 restatements (state dict in, tensors out; plain torch-CPU ops) of what the reference computes in
  * encode_video_with_siglip2   (hunyuanvideo_foley/utils/feature_utils.py:64-79)  -> siglip_pooler_output
  * encode_text_feat            (feature_utils.py:132-138)                         -> clap_last_hidden_state
  * encode_video_with_sync      (feature_utils.py:81-106)                          -> synchformer_visual
The first two algorithms live in a third-party dependency that is NOT under /root/reference: HF `transformers` (5.5.x in this
image) — models/siglip/modeling_siglip.py (SiglipVisionTransformer, SiglipMultiheadAttentionPoolingHead) and
models/clap/modeling_clap.py (ClapTextModel), called from the reference at nodes.py:199-201.  The third is the reference's own
models/synchformer/{motionformer,video_model_builder,vit_helper}.py with its divided_224_16x4.yaml.

Pinning (tests/test_oracle_golden.py): against the HF modules built from their configs on seeded weights (CPU, fp32, 1e-5), and
against the reference's own MotionFormer run through tools/ref_shims.py and the golden it wrote on a B200
(tests/golden/synchformer_d2.pt, fp32 output, 1e-4).  Arithmetic policy here is fp32 only: the engine's bf16 / fp16-autocast
rounding points are checked on the GPU against the modules themselves (tests/test_gpu_encoders.py), with this fp32 value as
the ground truth both sides are measured from.
"""
import math

import torch
import torch.nn.functional as F


def _ln(x, sd, name, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _mha(q, k, v, heads, mask=None):
    """softmax(q k^T / sqrt(d) [+ mask]) v over [B, S, heads * d] tensors; mask: [B, Sk] with 0 = masked key."""
    B, Sq, C = q.shape
    d = C // heads
    qh, kh, vh = (t.view(B, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Sq, C)


# ------------------------------------------------------------------------------------------------ SigLIP2 vision tower
def siglip_pooler_output(sd, pixels, heads=12, eps=1e-6, patch=16):
    """HF SiglipVisionTransformer.forward -> pooler_output (modeling_siglip.py: SiglipVisionEmbeddings.forward :175-186,
    SiglipEncoderLayer.forward :340-362, SiglipAttention.forward :275-312, SiglipMLP :323-327, post_layernorm + head
    :604-625, SiglipMultiheadAttentionPoolingHead.forward :639-654).  sd: `vision_model.*` keys; pixels [T, 3, H, W]."""
    p = "vision_model."
    x = F.conv2d(pixels, sd[p + "embeddings.patch_embedding.weight"], sd[p + "embeddings.patch_embedding.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2) + sd[p + "embeddings.position_embedding.weight"][None]
    i = 0
    while p + f"encoder.layers.{i}.layer_norm1.weight" in sd:
        l = p + f"encoder.layers.{i}."
        h = _ln(x, sd, l + "layer_norm1", eps)
        a = _mha(_lin(h, sd, l + "self_attn.q_proj"), _lin(h, sd, l + "self_attn.k_proj"), _lin(h, sd, l + "self_attn.v_proj"), heads)
        x = x + _lin(a, sd, l + "self_attn.out_proj")
        h = _ln(x, sd, l + "layer_norm2", eps)
        x = x + _lin(F.gelu(_lin(h, sd, l + "mlp.fc1"), approximate="tanh"), sd, l + "mlp.fc2")
        i += 1
    x = _ln(x, sd, p + "post_layernorm", eps)
    # attention pooling: one learned probe query per image (nn.MultiheadAttention with packed in_proj)
    C = x.shape[-1]
    W, b = sd[p + "head.attention.in_proj_weight"], sd[p + "head.attention.in_proj_bias"]
    probe = sd[p + "head.probe"].expand(x.shape[0], 1, C)
    q = F.linear(probe, W[:C], b[:C])
    k, v = F.linear(x, W[C:2 * C], b[C:2 * C]), F.linear(x, W[2 * C:], b[2 * C:])
    hs = _lin(_mha(q, k, v, heads), sd, p + "head.attention.out_proj")
    h = _ln(hs, sd, p + "head.layernorm", eps)
    hs = hs + _lin(F.gelu(_lin(h, sd, p + "head.mlp.fc1"), approximate="tanh"), sd, p + "head.mlp.fc2")
    return hs[:, 0]


# ------------------------------------------------------------------------------------------------ CLAP text tower
def clap_position_ids(input_ids, pad_id=1):
    """ClapTextEmbeddings.create_position_ids_from_input_ids (modeling_clap.py:1053-1067): non-pad tokens count up from
    pad_id + 1, pad tokens keep pad_id."""
    m = (input_ids != pad_id).long()
    return torch.cumsum(m, dim=1) * m + pad_id


def clap_last_hidden_state(sd, input_ids, attention_mask, heads=12, eps=1e-12, pad_id=1):
    """HF ClapTextModel.forward -> last_hidden_state (modeling_clap.py: ClapTextEmbeddings.forward :963-1010,
    ClapTextLayer :1203-1225 = self-attention with the key-padding mask + post-LN residual blocks, GELU-erf MLP).
    sd: `text_model.*` keys; padded QUERY rows are computed like any other (the reference hands them on)."""
    p = "text_model."
    x = (sd[p + "embeddings.word_embeddings.weight"][input_ids] + sd[p + "embeddings.token_type_embeddings.weight"][0]
         + sd[p + "embeddings.position_embeddings.weight"][clap_position_ids(input_ids, pad_id)])
    x = _ln(x, sd, p + "embeddings.LayerNorm", eps)
    i = 0
    while p + f"encoder.layer.{i}.attention.self.query.weight" in sd:
        l = p + f"encoder.layer.{i}."
        a = _mha(_lin(x, sd, l + "attention.self.query"), _lin(x, sd, l + "attention.self.key"), _lin(x, sd, l + "attention.self.value"),
                 heads, attention_mask)
        x = _ln(_lin(a, sd, l + "attention.output.dense") + x, sd, l + "attention.output.LayerNorm", eps)
        x = _ln(_lin(F.gelu(_lin(x, sd, l + "intermediate.dense")), sd, l + "output.dense") + x, sd, l + "output.LayerNorm", eps)
        i += 1
    return x


# ------------------------------------------------------------------------------------------------ Synchformer visual extractor
def _divided_attention(x, sd, name, heads, n_sp, n_t, mode):
    """DividedAttention.forward (vit_helper.py:56-114): the class token attends to every token; the patch tokens attend,
    per spatial location over time ("time", einops `b (f n) d -> (b n) f d`) or per frame over space ("space",
    `b (f n) d -> (b f) n d`), to [class token; their group]."""
    B, N, C = x.shape
    q, k, v = _lin(x, sd, name + ".qkv").chunk(3, dim=-1)
    cls_out = _mha(q[:, :1], k, v, heads)
    q_, k_, v_ = (t[:, 1:].view(B, n_t, n_sp, C) for t in (q, k, v))
    if mode == "time":
        q_, k_, v_ = (t.permute(0, 2, 1, 3).reshape(B * n_sp, n_t, C) for t in (q_, k_, v_))
        rep = n_sp
    else:
        q_, k_, v_ = (t.reshape(B * n_t, n_sp, C) for t in (q_, k_, v_))
        rep = n_t
    ck, cv = (t[:, :1].repeat_interleave(rep, dim=0) for t in (k, v))
    out = _mha(q_, torch.cat([ck, k_], 1), torch.cat([cv, v_], 1), heads)
    out = out.view(B, n_sp, n_t, C).permute(0, 2, 1, 3) if mode == "time" else out.view(B, n_t, n_sp, C)
    return _lin(torch.cat([cls_out, out.reshape(B, n_t * n_sp, C)], 1), sd, name + ".proj")


def synchformer_visual(sd, frames, heads=12, eps=1e-6):
    """encode_video_with_sync (feature_utils.py:81-106: 16-frame windows every 8 frames) -> Synchformer.forward
    (synchformer.py:43-50) -> MotionFormer.forward / forward_segments (motionformer.py:178-213) -> VisionTransformer
    .forward_features (video_model_builder.py:165-215: Conv3d tubelets, class token, "separate" position + temporal
    embeddings, DividedSpaceTimeBlock x depth: vit_helper.py:154-167) -> norm on the patch tokens -> spatial aggregation layer
    (motionformer.py:236-355: nn.TransformerEncoderLayer, norm_first, GELU, a class token per temporal position).
    sd: MotionFormer keys (or `vfeat_extractor.*`); frames [T, 3, 224, 224] -> [S * 8, C], S = (T - 16) // 8 + 1."""
    if any(k.startswith("vfeat_extractor.") for k in sd):
        sd = {k[len("vfeat_extractor."):]: v for k, v in sd.items() if k.startswith("vfeat_extractor.")}
    S = (frames.shape[0] - 16) // 8 + 1
    x = torch.stack([frames[i * 8: i * 8 + 16] for i in range(S)]).permute(0, 2, 1, 3, 4)        # [S, 3, 16, H, W]
    x = F.conv3d(x, sd["patch_embed_3d.proj.weight"], sd["patch_embed_3d.proj.bias"], stride=(2, 16, 16))
    C, n_t, n_sp = x.shape[1], x.shape[2], x.shape[3] * x.shape[4]
    x = x.flatten(2).transpose(1, 2)
    pos = sd["pos_embed"][:, 1:].repeat(1, n_t, 1) + sd["temp_embed"].repeat_interleave(n_sp, 1)
    x = torch.cat([sd["cls_token"].expand(S, -1, -1), x], 1) + torch.cat([sd["pos_embed"][:, :1], pos], 1)
    i = 0
    while f"blocks.{i}.norm1.weight" in sd:
        b = f"blocks.{i}."
        x = x + _divided_attention(_ln(x, sd, b + "norm3", eps), sd, b + "timeattn", heads, n_sp, n_t, "time")
        x = x + _divided_attention(_ln(x, sd, b + "norm1", eps), sd, b + "attn", heads, n_sp, n_t, "space")
        x = x + _lin(F.gelu(_lin(_ln(x, sd, b + "norm2", eps), sd, b + "mlp.fc1")), sd, b + "mlp.fc2")
        i += 1
    x = _ln(x[:, 1:], sd, "norm", eps)                                                            # [S, n_t * n_sp, C]
    a = "spatial_attn_agg."
    x = torch.cat([sd[a + "cls_token"].expand(S * n_t, -1, -1), x.reshape(S * n_t, n_sp, C)], 1)  # [(S t), 1 + n_sp, C]
    h = _ln(x, sd, a + "norm1", eps)
    W, bq = sd[a + "self_attn.in_proj_weight"], sd[a + "self_attn.in_proj_bias"]
    q, k, v = F.linear(h, W, bq).chunk(3, dim=-1)
    x = x + _lin(_mha(q, k, v, heads), sd, a + "self_attn.out_proj")
    x = x + _lin(F.gelu(_lin(_ln(x, sd, a + "norm2", eps), sd, a + "linear1")), sd, a + "linear2")
    return x[:, 0].reshape(S * n_t, C)
