"""Bridge to the condition encoders (SigLIP2, Synchformer, CLAP) — OUT OF THIS ENGINE'S SCOPE (SURVEY.md §8f:
they run once per clip, before the denoise path).  When the reference's model package `hunyuanvideo_foley`
and the HF checkpoints are available, this borrows them and exposes the single callable the Sampler needs:

    extract_features(frames_8fps|None, frames_25fps|None, prompt, negative_prompt)
        -> (visual_feats, text_feats, audio_len_in_s)

with the tensors the reference's `feature_process_from_tensors` returns (reference utils.py:262-292).
"""
import torch


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


def load_reference_extractors(synchformer_path, device, load_torch_file):
    """`load_torch_file`: comfy.utils.load_torch_file (reference nodes.py:177) — reads .safetensors and un-pickles
    .pth files safely; the caller (nodes.py) passes ComfyUI's, or its own weights_only fallback outside ComfyUI."""
    from transformers import AutoModel, AutoTokenizer, ClapTextModelWithProjection
    from hunyuanvideo_foley.models.synchformer import Synchformer          # reference package, if installed
    from hunyuanvideo_foley.utils.feature_utils import (encode_text_feat, encode_video_with_siglip2,
                                                        encode_video_with_sync)
    from .config import AttributeDict

    sd = load_torch_file(synchformer_path, device=torch.device("cpu"))
    if isinstance(sd, dict) and "state_dict" in sd and not any(torch.is_tensor(v) for v in sd.values()):
        sd = sd["state_dict"]
    sync_model = Synchformer()
    sync_model.load_state_dict(sd, strict=False)
    siglip2_preprocess, syncformer_preprocess = _v2_pipelines()
    deps = AttributeDict({
        "syncformer_model": sync_model.to(device).eval(),
        "siglip2_preprocess": siglip2_preprocess,
        "syncformer_preprocess": syncformer_preprocess,
        "siglip2_model": AutoModel.from_pretrained("google/siglip2-base-patch16-512").to(device).eval(),
        "clap_tokenizer": AutoTokenizer.from_pretrained("laion/larger_clap_general"),
        "clap_model": ClapTextModelWithProjection.from_pretrained("laion/larger_clap_general").to(device).eval(),
        "device": device,
    })
    def extract_features(pre_8fps, pre_25fps, prompt, negative_prompt):
        """pre_8fps [T8,3,512,512] / pre_25fps [T25,3,224,224]: encoder inputs already preprocessed on the GPU by
        preprocess.preprocess_video (the Sampler does that when `preprocessed_inputs` is set), or None for text-to-audio."""
        visual, audio_len = {}, None
        if pre_8fps is not None:
            visual["siglip2_feat"] = encode_video_with_siglip2(pre_8fps.unsqueeze(0).to(device), deps)
            visual["syncformer_feat"] = encode_video_with_sync(pre_25fps.unsqueeze(0).to(device), deps)
            audio_len = pre_25fps.shape[0] / 25.0
        feats, _ = encode_text_feat([negative_prompt, prompt], deps)
        return visual, {"text_feat": feats[1:], "uncond_text_feat": feats[:1]}, audio_len

    out = dict(deps)
    out["extract_features"] = extract_features
    out["preprocessed_inputs"] = True
    return out
