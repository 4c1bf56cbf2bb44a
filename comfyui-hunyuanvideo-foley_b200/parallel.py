"""Multi-GPU host logic: the batch-of-variations axis sharded across GPUs, in two forms:
  * one process per GPU under torchrun (bench.py --gpus N: broadcast_conditions / gather_waveforms over torch.distributed),
  * ONE process with one engine + one host thread per visible GPU (denoise_sharded: what the Sampler node uses when
    batch_size >= 2 and the process sees several GPUs — a ComfyUI server is one process).

The denoise path has no cross-sample operation (conditions are `.repeat`-ed per variation, reference
utils.py:159-162), so each rank runs its own fused cond+uncond batch with replicated weights and there is no
per-step communication.  The only exchanges are
  * ONE broadcast of the packed condition embeddings from rank 0 (NCCL over NVLink; gloo in CPU tests), and
  * ONE gather of the decoded waveforms to rank 0.
Noise is a single host draw for the global batch (seed parity with the 1-GPU run); rank r takes its rows.
"""
import os
import threading

import torch
import torch.distributed as dist

COND_KEYS = ("siglip2_feat", "syncformer_feat", "text_feat", "uncond_text_feat")


def shard_range(global_batch, world_size, rank):
    """Rows [lo, hi) of the global variation batch owned by `rank`: contiguous blocks, remainder to low ranks."""
    if global_batch < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad shard arguments")
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_conditions(feats, dtype=torch.bfloat16):
    """dict of [1, T_i, D] tensors -> (flat 1-D tensor, shapes) so that one collective moves all of them."""
    shapes = [tuple(feats[k].shape) for k in COND_KEYS]
    flat = torch.cat([feats[k].reshape(-1).to(dtype) for k in COND_KEYS])
    return flat, shapes


def unpack_conditions(flat, shapes):
    out, off = {}, 0
    for k, s in zip(COND_KEYS, shapes):
        n = 1
        for d in s:
            n *= d
        out[k] = flat[off:off + n].view(s)
        off += n
    return out


def broadcast_conditions(feats, shapes, device, src=0, dtype=torch.bfloat16, group=None):
    """Rank `src` passes its (host or device) condition dict; every rank returns the dict on `device`.
    `shapes` (list of 4 shapes) must be known on every rank — they follow from the clip duration."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = sum(int(torch.Size(s).numel()) for s in shapes)
    flat = torch.empty(n, dtype=dtype, device=device)
    if rank == src:
        packed, src_shapes = pack_conditions(feats, dtype)
        if [tuple(s) for s in src_shapes] != [tuple(s) for s in shapes]:
            raise ValueError(f"condition shapes {src_shapes} do not match the announced {shapes}")
        flat.copy_(packed, non_blocking=True)
    if world > 1:
        dist.broadcast(flat, src=src, group=group)
    return unpack_conditions(flat, shapes)


def gather_waveforms(local_wav, global_batch, dst=0, group=None):
    """[b_local, 1, T] on every rank -> [global_batch, 1, T] on rank `dst` (None elsewhere), rank order =
    variation order.  Ragged shards (global_batch not divisible) are padded to the largest shard for the
    collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_wav
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(global_batch, world, r) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = local_wav
    if local_wav.shape[0] < bmax:
        pad = torch.cat([local_wav, local_wav.new_zeros(bmax - local_wav.shape[0], *local_wav.shape[1:])])
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad.contiguous(), bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)])


# ------------------------------------------------------------------------------------------------ one process, N GPUs
def local_devices(batch_size, primary=None):
    """GPUs of THIS process a batch of `batch_size` variations is sharded over: the primary device first, then the other
    visible ones, never more devices than variations.  FOLEY_B200_GPUS = "auto" (default: all visible) or a count."""
    if not torch.cuda.is_available():
        return []
    primary = torch.device(primary) if primary is not None else torch.device("cuda", torch.cuda.current_device())
    if primary.type != "cuda":
        return []
    pi = primary.index if primary.index is not None else torch.cuda.current_device()
    env = os.environ.get("FOLEY_B200_GPUS", "auto").strip().lower()
    n_vis = torch.cuda.device_count()
    n = n_vis if env in ("", "auto") else max(1, min(int(env), n_vis))
    order = [pi] + [i for i in range(n_vis) if i != pi]
    return [torch.device("cuda", i) for i in order[:max(1, min(n, int(batch_size)))]]


def denoise_sharded(visual_feats, text_feats, audio_len_in_s, model_dict, cfg, guidance_scale, num_inference_steps,
                    batch_size, sampler, generator, devices):
    """`sampling.denoise_process_with_generator` for a batch of variations sharded over `devices` (all in this process):
    weights replicated (FoleyModel.on_device / FoleyDAC.on_device, cached), ONE broadcast of the packed condition
    embeddings from the primary GPU (torch.nn.parallel.comm.broadcast: NCCL when available, peer copies over NVLink otherwise),
    rank r runs rows shard_range(batch_size, n, r) of the one host noise draw on its own engine from its own host thread
    (the C ABI releases the GIL), ONE gather of the decoded waveforms to the primary GPU.  No per-step exchange.
    The result equals, bit for bit, running every shard alone on one GPU with `batch_slice`."""
    from torch.nn.parallel import comm   # single-process multi-GPU collectives (NCCL broadcast / peer-copy gather)

    from .config import AttributeDict
    from .sampling import denoise_process_with_generator, prepare_latents_with_generator
    devices = [torch.device(d) for d in devices]
    n, primary = len(devices), devices[0]
    model, dac = model_dict.foley_model, model_dict.dac_model
    models = [model.on_device(d) for d in devices]
    dacs = [dac.on_device(d) for d in devices]
    feats = {"siglip2_feat": visual_feats["siglip2_feat"][:1], "syncformer_feat": visual_feats["syncformer_feat"][:1],
             "text_feat": text_feats["text_feat"][:1], "uncond_text_feat": text_feats["uncond_text_feat"][:1]}
    flat, shapes = pack_conditions({k: v.to(primary) for k, v in feats.items()}, model.dtype)
    copies = comm.broadcast(flat, [d.index for d in devices])
    state = generator.get_state() if generator is not None else None
    results, errors = [None] * n, [None] * n

    def work(r):
        try:
            dev = devices[r]
            with torch.cuda.device(dev):
                f = unpack_conditions(copies[r], shapes)
                md = AttributeDict(dict(model_dict))
                md["foley_model"], md["dac_model"], md["device"] = models[r], dacs[r], dev
                md["report_progress"] = bool(model_dict.get("report_progress", True)) and r == 0   # one progress bar
                g = None
                if state is not None:
                    g = torch.Generator(device="cpu")
                    g.set_state(state)           # every shard slices the SAME host draw
                wav, sr = denoise_process_with_generator(
                    {"siglip2_feat": f["siglip2_feat"], "syncformer_feat": f["syncformer_feat"]},
                    {"text_feat": f["text_feat"], "uncond_text_feat": f["uncond_text_feat"]}, audio_len_in_s, md, cfg,
                    guidance_scale, num_inference_steps, batch_size, sampler, generator=g,
                    batch_slice=shard_range(batch_size, n, r))
                torch.cuda.current_stream(dev).synchronize()
                results[r] = (wav, sr)
        except BaseException as e:   # noqa: BLE001 - re-raised on the caller's thread
            errors[r] = e

    threads = [threading.Thread(target=work, args=(r,), name=f"foley-gpu{devices[r].index}") for r in range(1, n)]
    for t in threads:
        t.start()
    work(0)                                   # the primary shard (and the progress callbacks) on the caller's thread
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    if generator is not None:                 # leave the caller's generator where the 1-GPU path leaves it
        kw = cfg.model_config.model_kwargs
        prepare_latents_with_generator(None, batch_size, kw.audio_vae_latent_dim, int(audio_len_in_s * kw.audio_frame_rate),
                                       model.dtype, "cpu", generator)
    full = comm.gather([w for w, _ in results], dim=0, destination=primary.index)
    return full, results[0][1]
