"""ComfyUI node surface — the drop-in boundary (reference nodes.py:57-683).

The six node classes keep the reference's ids, display names, INPUT_TYPES / RETURN_TYPES / FUNCTION /
CATEGORY and socket types (HUNYUAN_MODEL, HUNYUAN_DEPS, TORCH_COMPILE_CFG, BLOCKSWAPARGS, IMAGE, AUDIO), so
saved workflows (example_workflows/HunyuanVideoFoleyExample.json) load and run unchanged.  What changes is
what flows through HUNYUAN_MODEL / dac_model: handles to the B200 engine instead of nn.Modules.

  * torch.compile and BlockSwap nodes are accepted and ignored (north-star: no torch.compile path, no
    block-swap / CPU-offload fallback; 10.3 GB of bf16 weights fit a B200's 180 GB).
  * The condition encoders (reference nodes.py:283-351) — SigLIP2, the Synchformer visual extractor and the CLAP text
    tower — run on the engine (encoders.py, csrc/encoders.cu; weights straight from the HF snapshots / the Synchformer
    checkpoint, tokenizer HF's; feature_bridge.py).  `deps` can also be populated by the caller with any callable
    `extract_features` (tests and bench use seeded stubs).
"""
import logging
import os

import torch

from .config import AttributeDict, detect_model_size, load_model_config
from .engine import FoleyEngine, FoleyError
from .sampling import denoise_process_with_generator

logger = logging.getLogger("HunyuanVideo-Foley-B200")

try:  # inside ComfyUI
    import folder_paths
    import comfy.model_management as mm
    from comfy.utils import load_torch_file
    foley_models_dir = os.path.join(folder_paths.models_dir, "foley")
    if "foley" not in folder_paths.folder_names_and_paths:   # reference nodes.py:25-27
        folder_paths.folder_names_and_paths["foley"] = ([foley_models_dir], folder_paths.supported_pt_extensions)
    _IN_COMFY = True
except Exception:  # outside ComfyUI (tests, bench, library use)
    folder_paths = None
    _IN_COMFY = False

    class mm:  # noqa: N801 - mirrors comfy.model_management
        @staticmethod
        def get_torch_device():
            return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

        @staticmethod
        def unet_offload_device():
            return torch.device("cpu")

        @staticmethod
        def soft_empty_cache():
            pass

    def load_torch_file(path, device=None):
        if str(path).endswith(".safetensors"):
            from safetensors.torch import load_file
            return load_file(path, device="cpu")
        obj = torch.load(path, map_location="cpu", weights_only=True)   # never un-pickle arbitrary objects
        return obj.get("state_dict", obj) if isinstance(obj, dict) else obj


def _check_precision(precision, resolved_dtype=None):
    """The reference computes in the dtype the loader picks (utils.py:222-239: fp32 without autocast, fp16 / bf16 under
    autocast).  The engine's arithmetic is bf16 tensor-core GEMMs with fp32 accumulation / residuals / norms: it is the
    reference's bf16 path.  Any other request is REFUSED rather than silently substituted, unless the user opts in with
    FOLEY_B200_PRECISION_FALLBACK=bf16 (then `precision` only sets the dtype of the noise draw and of the features, as in
    the reference, and the DiT still computes in bf16)."""
    want = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}.get(precision, resolved_dtype)
    if want in (None, torch.bfloat16):
        return
    if os.environ.get("FOLEY_B200_PRECISION_FALLBACK", "").lower() == "bf16":
        logger.warning("precision=%s requested: computing in bf16 (FOLEY_B200_PRECISION_FALLBACK=bf16)", precision)
        return
    raise FoleyError(f"precision={precision!r} (resolved {want}) is not implemented by the B200 engine, which computes the "
                     "reference's bf16 path (bf16 tensor-core GEMMs, fp32 accumulation, residuals and norms). Choose "
                     "precision='bf16', or set FOLEY_B200_PRECISION_FALLBACK=bf16 to accept bf16 arithmetic explicitly.")


def _foley_files(filter_substr=None):
    if folder_paths is None:
        return []
    files = folder_paths.get_filename_list("foley")
    return [f for f in files if filter_substr in f] if filter_substr else files


# --------------------------------------------------------------------------------------------------------
# Objects carried on the custom sockets
# --------------------------------------------------------------------------------------------------------
class FoleyModel:
    """HUNYUAN_MODEL: what the Sampler needs from the reference's HunyuanVideoFoley module (`.dtype`,
    `get_empty_*_sequence`, `.to`, the sticky `_text_len_fixed`) plus the engine handle."""

    def __init__(self, engine, empty_clip_feat, empty_sync_feat, cfg, dtype=torch.bfloat16):
        self.engine = engine
        self.cfg = cfg
        self.dtype = dtype
        self.empty_clip_feat = empty_clip_feat.detach().to(dtype)   # [1, clip_dim]  (hifi_foley.py:524-527)
        self.empty_sync_feat = empty_sync_feat.detach().to(dtype)
        kw = cfg.model_config.model_kwargs
        self.clip_len = kw.get("clip_length", 64)
        self.sync_len = kw.get("sync_length", 192)
        self._make_engine = None      # device -> FoleyEngine with the same weights (multi-GPU replicas, parallel.py)
        self._replicas = {}

    def on_device(self, device):
        """This model on another GPU of the process: the same weights loaded into a second engine (weights are replicated,
        variations are sharded: SURVEY.md §8e), built on first use and cached."""
        device = torch.device(device)
        if device == self.engine.device:
            return self
        if device.index not in self._replicas:
            if self._make_engine is None:
                raise FoleyError("this FoleyModel cannot be replicated (no weight source kept); build it with from_state_dict / from_safetensors")
            rep = FoleyModel(self._make_engine(device), self.empty_clip_feat, self.empty_sync_feat, self.cfg, self.dtype)
            logger.info("Replicated the HunyuanVideoFoley weights on %s", device)
            self._replicas[device.index] = rep
        return self._replicas[device.index]

    @classmethod
    def from_state_dict(cls, state_dict, model_size=None, device=None, dtype=torch.bfloat16, quantization="none"):
        """quantization: "none" or the FP8 storage mode the reference would wrap its Linear / Conv weights in
        (checkpoint.resolve_quantization); the engine rounds those weights through it and keeps computing in bf16."""
        model_size = model_size or detect_model_size(state_dict)
        cfg = load_model_config(model_size)
        engine = FoleyEngine(dict(cfg.model_config.model_kwargs), device=device, with_dac=False)
        def make(dev, sd=state_dict):
            e = FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev, with_dac=False)
            e.set_fp8_weight_storage(quantization)
            e.load_state_dict(sd)
            e.finalize()
            return e
        m = cls(make(device), state_dict["empty_clip_feat"], state_dict["empty_sync_feat"], cfg, dtype)
        if torch.cuda.is_available() and torch.cuda.device_count() > 1:
            m._make_engine = make          # (keeps the state dict alive: only when there is a second GPU to replicate onto)
        return m

    @classmethod
    def from_safetensors(cls, path, precision="bf16", quantization="auto", device=None, cfg=None):
        """The Model Loader's work (reference nodes.py:72-126) without a host state dict: header-only dtype sniffing
        (utils.py:492-515), then file -> device through foley_engine_load_safetensors."""
        from . import checkpoint as ck
        _check_precision(precision)
        header, data_start = ck.read_header(path)
        dts = [(ck.ST_DTYPES.get(e["dtype"]), int(torch.tensor(e["shape"]).prod()) if e["shape"] else 1)
               for e in header.values()]
        dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}.get(precision, torch.bfloat16)
        if precision == "auto":
            dtype = ck.detect_major_precision(dts)
            logger.info("Auto precision selected from checkpoint: %s", dtype)
            _check_precision(precision, dtype)
        qmode = ck.resolve_quantization(quantization, ck.detect_fp8(d for d, _ in dts))
        w = header.get("audio_embedder.proj.weight")
        model_size = "xl" if w is not None and int(w["shape"][0]) == 1408 else "xxl"
        cfg = cfg if cfg is not None else load_model_config(model_size)
        def make(dev):
            e = FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev, with_dac=False)
            e.set_fp8_weight_storage(qmode)
            n_ = e.load_safetensors(path)
            e.finalize()
            logger.info("Loaded %d tensors from %s straight into the B200 engine on %s (fp8 storage emulation: %s)", n_, path, e.device, qmode)
            return e
        m = cls(make(device), ck.read_tensor(path, "empty_clip_feat", header, data_start),
                ck.read_tensor(path, "empty_sync_feat", header, data_start), cfg, dtype)
        m._make_engine = make              # replicas re-read the file
        return m

    def get_empty_clip_sequence(self, bs=None, len=None):      # hifi_foley.py:620-625
        len = len if len is not None else self.clip_len
        if bs is None:
            return self.empty_clip_feat.expand(len, -1)
        return self.empty_clip_feat.unsqueeze(0).expand(bs, len, -1)

    def get_empty_sync_sequence(self, bs=None, len=None):      # hifi_foley.py:627-632
        len = len if len is not None else self.sync_len
        if bs is None:
            return self.empty_sync_feat.expand(len, -1)
        return self.empty_sync_feat.unsqueeze(0).expand(bs, len, -1)

    def to(self, *args, **kwargs):   # weights stay resident on the B200; device moves are no-ops
        return self

    def eval(self):
        return self

    def block_swap(self, **kwargs):  # accepted and ignored
        logger.info("BlockSwap settings ignored: the B200 engine keeps all blocks resident")


class FoleyDAC:
    """`deps['dac_model']`: DAC-VAE decoder on the engine (`decode`, `sample_rate`, `hop_length`)."""

    sample_rate = 48000
    hop_length = 960

    def __init__(self, engine):
        self.engine = engine
        self._make_engine = None
        self._replicas = {}

    @classmethod
    def from_state_dict(cls, state_dict, device=None):
        # DAC-only engine instance; the DiT dimensions are irrelevant for decode
        cfg = load_model_config("xxl")

        def make(dev, sd=state_dict):
            e = FoleyEngine(dict(cfg.model_config.model_kwargs), device=dev, with_dac=False)
            e.load_state_dict(sd, prefix="dac.")
            return e
        d = cls(make(device))
        if torch.cuda.is_available() and torch.cuda.device_count() > 1:
            d._make_engine = make          # the decoder is small (~300 MB fp32): the state dict is kept for replicas
        return d

    def on_device(self, device):
        """The decoder on another GPU of the process (multi-GPU Sampler, parallel.py)."""
        device = torch.device(device)
        if device == self.engine.device:
            return self
        if device.index not in self._replicas:
            if self._make_engine is None:
                raise FoleyError("this FoleyDAC cannot be replicated (no weight source kept)")
            self._replicas[device.index] = FoleyDAC(self._make_engine(device))
        return self._replicas[device.index]

    def decode(self, z):
        from . import torch_ops as ops
        return ops.dac_decode(self.engine, z.to(self.engine.device))

    def to(self, *args, **kwargs):
        return self

    def parameters(self):
        return iter(())


# --------------------------------------------------------------------------------------------------------
# NODE 1: Hunyuan Model Loader (reference nodes.py:57-151)
# --------------------------------------------------------------------------------------------------------
class HunyuanModelLoader:
    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {
                "model_name": (_foley_files(),),
                "precision": (["auto", "bf16", "fp16", "fp32"], {"default": "bf16", "tooltip": "The B200 engine computes the reference's bf16 path; fp16 / fp32 are refused unless FOLEY_B200_PRECISION_FALLBACK=bf16 is set"}),
                "quantization": (["none", "fp8_e4m3fn", "fp8_e5m2", "auto"], {"default": "auto", "tooltip": "Same numbers as the reference's FP8 weight-only storage (weights rounded through FP8 at load); the engine keeps them resident in bf16"}),
            },
        }

    RETURN_TYPES = ("HUNYUAN_MODEL",)
    FUNCTION = "build_model"
    CATEGORY = "audio/HunyuanFoley"

    def load_model(self, model_name, precision, quantization):
        if folder_paths is None:
            raise FoleyError("HunyuanModelLoader needs ComfyUI's folder_paths; use FoleyModel.from_state_dict outside ComfyUI")
        model_path = folder_paths.get_full_path("foley", model_name)
        if model_path is None or not os.path.exists(model_path):
            raise FileNotFoundError(f"Hunyuan-Foley checkpoint not found: {model_name}")
        _check_precision(precision)
        if str(model_path).endswith(".safetensors"):
            model = FoleyModel.from_safetensors(model_path, precision, quantization, device=mm.get_torch_device())
        else:   # .pt / .pth checkpoints go through a host state dict like the reference
            from . import checkpoint as ck
            state_dict = load_torch_file(model_path, device=mm.unet_offload_device())
            tensors = [v for v in state_dict.values() if isinstance(v, torch.Tensor)]
            dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}.get(precision, torch.bfloat16)
            if precision == "auto":
                dtype = ck.detect_major_precision((v.dtype, v.numel()) for v in tensors)
                _check_precision(precision, dtype)
            qmode = ck.resolve_quantization(quantization, ck.detect_fp8(v.dtype for v in tensors))
            model = FoleyModel.from_state_dict(state_dict, device=mm.get_torch_device(), dtype=dtype, quantization=qmode)
            del state_dict
        logger.info("Loaded HunyuanVideoFoley main model into the B200 engine: %s", model_name)
        return model

    def build_model(self, model_name, precision, quantization):
        return (self.load_model(model_name, precision, quantization),)


# --------------------------------------------------------------------------------------------------------
# NODE 2: Hunyuan Dependencies Loader (reference nodes.py:156-206)
# --------------------------------------------------------------------------------------------------------
class HunyuanDependenciesLoader:
    @classmethod
    def INPUT_TYPES(cls):
        return {"required": {"vae_name": (_foley_files("vae"),), "synchformer_name": (_foley_files("synch"),)}}

    RETURN_TYPES = ("HUNYUAN_DEPS",)
    FUNCTION = "load_dependencies"
    CATEGORY = "audio/HunyuanFoley"

    def load_dependencies(self, vae_name, synchformer_name):
        if folder_paths is None:
            raise FoleyError("HunyuanDependenciesLoader needs ComfyUI's folder_paths")
        device = mm.get_torch_device()
        deps = {}
        vae_sd = load_torch_file(folder_paths.get_full_path("foley", vae_name), device=torch.device("cpu"))
        if isinstance(vae_sd, dict) and "state_dict" in vae_sd:
            vae_sd = vae_sd["state_dict"]
        deps["dac_model"] = FoleyDAC.from_state_dict(vae_sd, device=device)
        # Condition encoders: SigLIP2, Synchformer and CLAP text all run on the engine (encoders.py); the tokenizer is HF's.
        try:
            from .feature_bridge import load_extractors
            deps.update(load_extractors(folder_paths.get_full_path("foley", synchformer_name), device, load_torch_file))
        except FoleyError:
            raise
        except Exception as e:  # noqa: BLE001
            raise FoleyError("loading the condition encoders failed: the SigLIP2 / CLAP weights and tokenizer come from their "
                             f"HF snapshots (transformers), the Synchformer weights from the selected checkpoint ({e})") from e
        deps["device"] = device
        return (AttributeDict(deps),)


# --------------------------------------------------------------------------------------------------------
# NODE 3: Hunyuan Foley Sampler (reference nodes.py:211-427)
# --------------------------------------------------------------------------------------------------------
def resample_frame_indices(num_frames_to_process, duration, fps):
    """torch.linspace(0, n-1, int(duration*fps)).long() — the reference's frame pick (nodes.py:310,315)."""
    return torch.linspace(0, num_frames_to_process - 1, int(duration * fps)).long()   # (same rule: preprocess.py)


def t2a_feature_lengths(duration):
    """Clip / sync token counts for text-to-audio (nodes.py:326-331)."""
    clip_seq_len = int(duration * 8)
    num_sync_frames = int(duration * 25)
    num_sync_segments = (num_sync_frames - 16) // 8 + 1
    return clip_seq_len, int(num_sync_segments * 8)


class HunyuanFoleySampler:
    SAMPLER_NAMES = ["euler", "heun-2", "midpoint-2", "kutta-4"]

    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {
                "hunyuan_model": ("HUNYUAN_MODEL",),
                "hunyuan_deps": ("HUNYUAN_DEPS",),
                "frame_rate": ("FLOAT", {"default": 16, "min": 1, "max": 120, "step": 0.1, "tooltip": "The framerate of the input image sequence"}),
                "duration": ("FLOAT", {"default": 5.0, "min": 1, "max": 60.0, "step": 0.1, "tooltip": "Duration of the audio to generate in seconds"}),
                "prompt": ("STRING", {"multiline": True, "default": "A person walks on frozen ice"}),
                "negative_prompt": ("STRING", {"multiline": True, "default": "noisy, harsh"}),
                "cfg_scale": ("FLOAT", {"default": 4.5, "min": 1.0, "max": 10.0, "step": 0.1, "tooltip": "Classifier-Free Guidance scale"}),
                "steps": ("INT", {"default": 50, "min": 10, "max": 100, "step": 1, "tooltip": "Number of denoising steps"}),
                "sampler": (cls.SAMPLER_NAMES, {"default": "euler", "tooltip": "Flow-matching ODE solver; euler runs fully inside the B200 engine (one CUDA graph per step)"}),
                "batch_size": ("INT", {"default": 1, "min": 1, "max": 64, "step": 1, "tooltip": "Number of audio variations to generate at once"}),
                "seed": ("INT", {"default": 0, "min": 0, "max": 0xffffffffffffffff}),
                "force_offload": ("BOOLEAN", {"default": True, "tooltip": "Accepted for compatibility; the engine keeps weights resident"}),
            },
            "optional": {
                "image": ("IMAGE",),
                "torch_compile_cfg": ("TORCH_COMPILE_CFG", {"tooltip": "Accepted and ignored (no torch.compile path)."}),
                "block_swap_args": ("BLOCKSWAPARGS", {"tooltip": "Accepted and ignored (no CPU offload path)."}),
            },
        }

    RETURN_TYPES = ("AUDIO", "AUDIO")
    RETURN_NAMES = ("audio_first", "audio_batch")
    FUNCTION = "generate_audio"
    CATEGORY = "audio/HunyuanFoley"

    def generate_audio(self, hunyuan_model, hunyuan_deps, frame_rate, duration, prompt, negative_prompt, cfg_scale,
                       steps, sampler, batch_size, seed, force_offload, image=None, torch_compile_cfg=None,
                       block_swap_args=None):
        device = mm.get_torch_device()
        hunyuan_cfg = hunyuan_model.cfg                      # the reference re-reads the XXL YAML here (nodes.py:269-271)
        rng = torch.Generator(device="cpu").manual_seed(seed)
        target_dtype = hunyuan_model.dtype
        if torch_compile_cfg is not None:
            logger.info("torch_compile_cfg ignored: the step already runs as one CUDA graph of hand-written kernels")
        if block_swap_args is not None:
            hunyuan_model.block_swap(**block_swap_args)

        # ---- Phase 1: condition features (reference nodes.py:283-351)
        visual_feats, audio_len_in_s = {}, duration
        extract = hunyuan_deps.get("extract_features") if isinstance(hunyuan_deps, dict) else None
        if extract is None:
            raise FoleyError("hunyuan_deps carries no feature extractors (see HunyuanDependenciesLoader)")
        if image is not None and hunyuan_deps.get("preprocessed_inputs", False):
            # frames -> encoder inputs on the GPU (quantise, 8 / 25 fps picks, antialiased bicubic resize, crop,
            # normalise: bit-exact with the reference's per-frame torchvision pipelines on the CPU)
            from .preprocess import preprocess_video
            pre8, pre25, _ = preprocess_video(image, duration, frame_rate, device)
            visual_feats, text_feats, audio_len_in_s = extract(pre8, pre25, prompt, negative_prompt)
        elif image is not None:
            total_input_frames = image.shape[0]
            num_frames_to_process = int(duration * frame_rate)
            if num_frames_to_process > total_input_frames:   # hold the last frame (nodes.py:298-303)
                pad = image[-1:].repeat(num_frames_to_process - total_input_frames, 1, 1, 1)
                image_slice = torch.cat((image, pad), dim=0)
            else:
                image_slice = image[:num_frames_to_process]
            image_slice = (image_slice * 255.0).byte().permute(0, 3, 1, 2)
            frames_8fps = image_slice.index_select(0, resample_frame_indices(num_frames_to_process, duration, 8))
            frames_25fps = image_slice.index_select(0, resample_frame_indices(num_frames_to_process, duration, 25))
            visual_feats, text_feats, audio_len_in_s = extract(frames_8fps, frames_25fps, prompt, negative_prompt)
        else:
            clip_seq_len, sync_seq_len = t2a_feature_lengths(duration)
            visual_feats["siglip2_feat"] = hunyuan_model.get_empty_clip_sequence(bs=1, len=clip_seq_len).to("cpu", dtype=target_dtype)
            visual_feats["syncformer_feat"] = hunyuan_model.get_empty_sync_sequence(bs=1, len=sync_seq_len).to("cpu", dtype=target_dtype)
            _, text_feats, _ = extract(None, None, prompt, negative_prompt)

        # ---- Phase 2: denoise + decode on the engine (reference nodes.py:353-407)
        model_dict = AttributeDict(dict(hunyuan_deps))
        model_dict["foley_model"] = hunyuan_model
        model_dict["device"] = device
        # batch_size is the only parallel axis of the path (reference nodes.py:228): with more than one visible GPU the
        # variations are sharded over them — one engine + one host thread per GPU inside this process, ONE broadcast of the
        # condition embeddings, ONE gather of the decoded waveforms (parallel.py; FOLEY_B200_GPUS=1 keeps one GPU)
        from . import parallel
        devices = parallel.local_devices(batch_size, device)
        if len(devices) > 1:
            decoded_waveform, sample_rate = parallel.denoise_sharded(
                visual_feats, text_feats, audio_len_in_s, model_dict, hunyuan_cfg, guidance_scale=cfg_scale,
                num_inference_steps=steps, batch_size=batch_size, sampler=sampler, generator=rng, devices=devices)
        else:
            decoded_waveform, sample_rate = denoise_process_with_generator(
                visual_feats, text_feats, audio_len_in_s, model_dict, hunyuan_cfg, guidance_scale=cfg_scale,
                num_inference_steps=steps, batch_size=batch_size, sampler=sampler, generator=rng)
        from .sampling import to_host
        waveform_batch = to_host(decoded_waveform.float())
        audio_output_first = {"waveform": waveform_batch[0].unsqueeze(0), "sample_rate": sample_rate}
        audio_output_batch = {"waveform": waveform_batch, "sample_rate": sample_rate}
        return (audio_output_first, audio_output_batch)


# --------------------------------------------------------------------------------------------------------
# NODE: Torch Compile settings (reference nodes.py:433-607) — socket kept, config ignored by the Sampler
# --------------------------------------------------------------------------------------------------------
class HunyuanFoleyTorchCompile:
    DESCRIPTION = "Kept for workflow compatibility. The B200 engine does not use torch.compile."

    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {
                "backend": (["inductor"], {"default": "inductor"}),
                "fullgraph": ("BOOLEAN", {"default": False}),
                "mode": (["default", "reduce-overhead", "max-autotune"], {"default": "default"}),
                "dynamic": (["true", "false", "None"], {"default": "false"}),
                "dynamo_cache_limit": ("INT", {"default": 64, "min": 64, "max": 8192, "step": 64}),
            }
        }

    RETURN_TYPES = ("TORCH_COMPILE_CFG",)
    FUNCTION = "make_config"
    CATEGORY = "audio/HunyuanFoley"

    def make_config(self, backend, mode, dynamic, fullgraph, dynamo_cache_limit):
        dyn_map = {"true": True, "false": False, "None": None}
        return ({"backend": backend, "mode": mode, "dynamic": dyn_map.get(str(dynamic), False),
                 "fullgraph": fullgraph, "dynamo_cache_limit": int(dynamo_cache_limit)},)


# --------------------------------------------------------------------------------------------------------
# NODE: BlockSwap settings (reference nodes.py:609-631) — socket kept, settings ignored
# --------------------------------------------------------------------------------------------------------
class HunyuanBlockSwap:
    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {"blocks_to_swap": ("INT", {"default": 30, "min": 0, "max": 57, "step": 1})},
            "optional": {
                "use_non_blocking": ("BOOLEAN", {"default": False}),
                "prefetch_blocks": ("INT", {"default": 1, "min": 0, "max": 10, "step": 1}),
                "block_swap_debug": ("BOOLEAN", {"default": False}),
            },
        }

    RETURN_TYPES = ("BLOCKSWAPARGS",)
    RETURN_NAMES = ("block_swap_args",)
    FUNCTION = "set_args"
    CATEGORY = "audio/HunyuanFoley"
    DESCRIPTION = "Kept for workflow compatibility. The B200 engine keeps all transformer blocks resident."

    def set_args(self, **kwargs):
        return (kwargs,)


# --------------------------------------------------------------------------------------------------------
# HELPER NODE: Select Audio From Batch (reference nodes.py:636-663)
# --------------------------------------------------------------------------------------------------------
class SelectAudioFromBatch:
    @classmethod
    def INPUT_TYPES(cls):
        return {
            "required": {
                "audio_batch": ("AUDIO", {"tooltip": "An audio object containing a batch of waveforms."}),
                "index": ("INT", {"default": 0, "min": 0, "max": 63, "tooltip": "The 0-based index of the audio to select from the batch."}),
            }
        }

    RETURN_TYPES = ("AUDIO",)
    FUNCTION = "select_audio"
    CATEGORY = "audio/utils"

    def select_audio(self, audio_batch, index):
        waveform_batch = audio_batch["waveform"]
        if index >= waveform_batch.shape[0]:
            logger.warning("Index %d is out of bounds for audio batch of size %d. Clamping to last item.", index,
                           waveform_batch.shape[0])
            index = waveform_batch.shape[0] - 1
        return ({"waveform": waveform_batch[index].unsqueeze(0), "sample_rate": audio_batch["sample_rate"]},)


NODE_CLASS_MAPPINGS = {
    "HunyuanModelLoader": HunyuanModelLoader,
    "HunyuanDependenciesLoader": HunyuanDependenciesLoader,
    "HunyuanFoleySampler": HunyuanFoleySampler,
    "HunyuanFoleyTorchCompile": HunyuanFoleyTorchCompile,
    "HunyuanBlockSwap": HunyuanBlockSwap,
    "SelectAudioFromBatch": SelectAudioFromBatch,
}
NODE_DISPLAY_NAME_MAPPINGS = {
    "HunyuanModelLoader": "Hunyuan-Foley Model Loader",
    "HunyuanDependenciesLoader": "Hunyuan-Foley Dependencies Loader",
    "HunyuanFoleySampler": "Hunyuan-Foley Sampler",
    "HunyuanFoleyTorchCompile": "Hunyuan-Foley Torch Compile",
    "HunyuanBlockSwap": "Hunyuan-Foley BlockSwap Settings",
    "SelectAudioFromBatch": "Select Audio From Batch",
}
