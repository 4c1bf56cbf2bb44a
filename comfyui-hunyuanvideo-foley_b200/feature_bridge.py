"""The condition encoders behind the Sampler (reference nodes.py:283-351, utils.py:262-292, feature_utils.py:64-138).

All three run ON THE ENGINE (encoders.py -> csrc/encoders.cu): SigLIP2 (google/siglip2-base-patch16-512) and the CLAP text
tower (laion/larger_clap_general) take their weights from the HF snapshot's model.safetensors straight to the device, the
Synchformer visual extractor (MotionFormer under fp16 autocast, feature_utils.py:81-106) from the checkpoint the
Dependencies Loader names; no torch module is built and the reference's model package is not needed.  The tokenizer is HF's
(host-side string work).

`load_extractors(...)` returns the `deps` entries the Dependencies Loader publishes (same keys as reference
nodes.py:176-201) plus `extract_features(pre_8fps, pre_25fps, prompt, negative_prompt) -> (visual_feats, text_feats,
audio_len_in_s)`, the tensors of the reference's `feature_process_from_tensors`.
"""
import logging

import torch

from .engine import FoleyError

logger = logging.getLogger("foley_b200")

SIGLIP2_REPO = "google/siglip2-base-patch16-512"      # reference nodes.py:199
CLAP_REPO = "laion/larger_clap_general"               # reference nodes.py:200-201


def _v2_pipelines():
    """The two torchvision v2 pipelines the reference's Dependencies Loader publishes in `deps` (nodes.py:184-196).
    The Sampler of this pack preprocesses on the GPU (preprocess.py, bit-exact with these); the objects are kept in
    `deps` because other nodes of a saved workflow may read the keys."""
    from torchvision.transforms import v2
    siglip2 = v2.Compose([v2.Resize((512, 512), interpolation=v2.InterpolationMode.BICUBIC, antialias=True),
                          v2.ToDtype(torch.float32, scale=True), v2.Normalize(mean=[0.5] * 3, std=[0.5] * 3)])
    sync = v2.Compose([v2.Resize(224, interpolation=v2.InterpolationMode.BICUBIC, antialias=True), v2.CenterCrop(224),
                       v2.ToDtype(torch.float32, scale=True), v2.Normalize(mean=[0.5] * 3, std=[0.5] * 3)])
    return siglip2, sync


def _snapshot_file(repo, filename):
    """Path of a file of the HF snapshot (downloaded on first use like `from_pretrained` does), or None."""
    try:
        from transformers.utils import cached_file
        return cached_file(repo, filename)
    except Exception as e:  # noqa: BLE001 — offline without a cache, or the repo ships .bin weights only
        logger.info("no %s in the snapshot of %s (%s)", filename, repo, e)
        return None


def load_siglip2(device, repo=SIGLIP2_REPO):
    """SigLIP2 vision tower + pooling head on the engine: config from the snapshot, weights file -> device."""
    from transformers import AutoConfig
    from .encoders import SiglipVisionEncoder
    vc = AutoConfig.from_pretrained(repo)
    vc = getattr(vc, "vision_config", vc)
    cfg = dict(hidden_size=vc.hidden_size, num_heads=vc.num_attention_heads, num_layers=vc.num_hidden_layers,
               intermediate_size=vc.intermediate_size, layer_norm_eps=vc.layer_norm_eps, image_size=vc.image_size,
               patch_size=vc.patch_size)
    path = _snapshot_file(repo, "model.safetensors")
    if path is not None:
        enc = SiglipVisionEncoder(cfg, device=device)
        enc.load_safetensors(path)
        return enc.finalize()
    from transformers import AutoModel
    return SiglipVisionEncoder.from_hf(AutoModel.from_pretrained(repo), device=device)


def load_clap_text(device, repo=CLAP_REPO):
    """(tokenizer, CLAP text tower on the engine)."""
    from transformers import AutoConfig, AutoTokenizer
    from .encoders import ClapTextEncoder
    tokenizer = AutoTokenizer.from_pretrained(repo)
    tc = AutoConfig.from_pretrained(repo)
    tc = getattr(tc, "text_config", tc)
    cfg = dict(hidden_size=tc.hidden_size, num_heads=tc.num_attention_heads, num_layers=tc.num_hidden_layers,
               intermediate_size=tc.intermediate_size, layer_norm_eps=tc.layer_norm_eps, vocab_size=tc.vocab_size,
               max_positions=tc.max_position_embeddings, pad_token_id=tc.pad_token_id)
    path = _snapshot_file(repo, "model.safetensors")
    if path is not None:
        enc = ClapTextEncoder(cfg, device=device)
        enc.load_safetensors(path)
        return tokenizer, enc.finalize()
    from transformers import ClapTextModelWithProjection
    return tokenizer, ClapTextEncoder.from_hf(ClapTextModelWithProjection.from_pretrained(repo), device=device)


def load_synchformer(synchformer_path, device, load_torch_file):
    """The Synchformer visual extractor on the engine, from the reference's checkpoint (nodes.py:176-181: `load_torch_file`
    + `load_state_dict(strict=False)`; keys `vfeat_extractor.*`, everything else is ignored by name)."""
    from .encoders import SynchformerEncoder
    sd = load_torch_file(synchformer_path, device=torch.device("cpu"))
    if isinstance(sd, dict) and "state_dict" in sd and not any(torch.is_tensor(v) for v in sd.values()):
        sd = sd["state_dict"]
    return SynchformerEncoder.from_state_dict(sd, device=device)


def make_extract_features(siglip2, clap_tokenizer, clap_text, sync_encode, device, max_text_tokens=None):
    """The one callable the Sampler needs.  siglip2 / clap_text: objects with `.encode` (encoders.py);
    sync_encode(frames [1, T25, 3, 224, 224]) -> [1, S, 768], or a string saying why no Synchformer is available."""
    def extract_features(pre_8fps, pre_25fps, prompt, negative_prompt):
        """pre_8fps [T8,3,512,512] / pre_25fps [T25,3,224,224]: encoder inputs preprocessed on the GPU
        (preprocess.preprocess_video), or None for text-to-audio."""
        visual, audio_len = {}, None
        if pre_8fps is not None:
            if isinstance(sync_encode, str):
                raise FoleyError(f"video-to-audio needs the Synchformer encoder: {sync_encode}")
            visual["siglip2_feat"] = siglip2.encode(pre_8fps.to(device)).unsqueeze(0)        # [1, T8, 768] (feature_utils.py:77-78)
            visual["syncformer_feat"] = sync_encode(pre_25fps.unsqueeze(0).to(device))
            audio_len = pre_25fps.shape[0] / 25.0                                             # utils.py:281
        tok = clap_tokenizer([negative_prompt, prompt], padding=True, return_tensors="pt")    # feature_utils.py:134, utils.py:284
        ids, mask = tok["input_ids"], tok["attention_mask"]
        if max_text_tokens is not None and ids.shape[1] > max_text_tokens:
            raise FoleyError(f"prompt has {ids.shape[1]} tokens, the text encoder takes {max_text_tokens}")
        feats = clap_text.encode(ids, mask)
        return visual, {"text_feat": feats[1:], "uncond_text_feat": feats[:1]}, audio_len
    return extract_features


def load_extractors(synchformer_path, device, load_torch_file):
    """`load_torch_file`: comfy.utils.load_torch_file (reference nodes.py:177) — reads .safetensors and un-pickles
    .pth files safely; the caller (nodes.py) passes ComfyUI's, or its own weights_only fallback outside ComfyUI."""
    from .config import AttributeDict
    siglip2 = load_siglip2(device)
    clap_tokenizer, clap_text = load_clap_text(device)
    sync = load_synchformer(synchformer_path, device, load_torch_file)
    siglip2_preprocess, syncformer_preprocess = _v2_pipelines()
    deps = AttributeDict({
        "syncformer_model": sync,              # engine encoders (.encode), not torch modules
        "siglip2_preprocess": siglip2_preprocess,
        "syncformer_preprocess": syncformer_preprocess,
        "siglip2_model": siglip2,
        "clap_tokenizer": clap_tokenizer,
        "clap_model": clap_text,
        "device": device,
    })

    def sync_encode(frames):                   # [1, T25, 3, 224, 224] -> [1, S * 8, 768]  (encode_video_with_sync)
        return sync.encode(frames[0])

    out = dict(deps)
    out["extract_features"] = make_extract_features(siglip2, clap_tokenizer, clap_text, sync_encode, device,
                                                    max_text_tokens=clap_text.config["max_positions"] - 2)
    out["preprocessed_inputs"] = True
    return out


# Name kept for callers of the round-1 bridge.
load_reference_extractors = load_extractors
