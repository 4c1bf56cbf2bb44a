"""Multi-GPU host logic: one process per GPU, the batch-of-variations axis sharded across ranks.

The denoise path has no cross-sample operation (conditions are `.repeat`-ed per variation, reference
utils.py:159-162), so each rank runs its own fused cond+uncond batch with replicated weights and there is no
per-step communication.  The only exchanges are
  * ONE broadcast of the packed condition embeddings from rank 0 (NCCL over NVLink; gloo in CPU tests), and
  * ONE gather of the decoded waveforms to rank 0.
Noise is a single host draw for the global batch (seed parity with the 1-GPU run); rank r takes its rows.
"""
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
