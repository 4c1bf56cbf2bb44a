"""Checkpoint inspection for the Model Loader (reference nodes.py:72-126, utils.py:492-515) without materialising a
host state dict: the .safetensors header tells dtypes and shapes, the engine maps the file and copies tensors straight
to the device (foley_engine_load_safetensors); only the two tiny `empty_*_feat` vectors the Sampler needs on the host
are read here.
"""
import json
import struct

import torch

ST_DTYPES = {"BF16": torch.bfloat16, "F32": torch.float32, "F16": torch.float16, "F64": torch.float64,
             "F8_E4M3": torch.float8_e4m3fn, "F8_E5M2": torch.float8_e5m2, "I64": torch.int64, "I32": torch.int32,
             "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool}
_ST_NAMES = {v: k for k, v in ST_DTYPES.items()}


def read_header(path):
    """{name: {"dtype", "shape", "data_offsets"}} and the byte offset of the data section."""
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(n))
    header.pop("__metadata__", None)
    return header, 8 + n


def read_tensor(path, name, header=None, data_start=None):
    """One tensor by name (used for empty_clip_feat / empty_sync_feat, a few KB)."""
    if header is None:
        header, data_start = read_header(path)
    e = header[name]
    lo, hi = e["data_offsets"]
    with open(path, "rb") as f:
        f.seek(data_start + lo)
        buf = bytearray(f.read(hi - lo))
    return torch.frombuffer(buf, dtype=ST_DTYPES[e["dtype"]]).reshape(e["shape"]).clone()


def write_safetensors(path, tensors, metadata=None):
    """Minimal writer (tests, tools): same container the reference's checkpoints use."""
    header, blobs, off = {}, [], 0
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    for k, t in tensors.items():
        t = t.detach().contiguous().cpu()
        raw = t.view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[k] = {"dtype": _ST_NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def detect_fp8(dtypes):
    """utils.py:492-503 over an iterable of torch dtypes in checkpoint order: first FP8 dtype seen, else None."""
    for dt in dtypes:
        if dt == torch.float8_e5m2:
            return "fp8_e5m2"
        if dt == torch.float8_e4m3fn:
            return "fp8_e4m3fn"
    return None


def detect_major_precision(dtype_numels):
    """utils.py:506-515 over (dtype, numel) pairs: the dtype among {bf16, fp16, fp32} holding most elements."""
    counts = {torch.bfloat16: 0, torch.float16: 0, torch.float32: 0}
    for dt, n in dtype_numels:
        if dt in counts:
            counts[dt] += int(n)
    if all(c == 0 for c in counts.values()):
        return torch.bfloat16
    return max(counts, key=counts.get)


def resolve_quantization(quantization, detected_fp8, capability_major=10):
    """nodes.py:106-121: the FP8 storage mode the reference would use, or None for "none"."""
    if quantization == "none":
        return None
    if quantization == "auto":
        if capability_major < 9:
            return "fp8_e5m2"
        return detected_fp8 if detected_fp8 is not None else "fp8_e4m3fn"
    return quantization


def round_through_fp8(t, mode):
    """Host statement of what "fp8_weight_storage" does on the device: value -> bf16 parameter -> FP8 buffer -> upcast."""
    qd = torch.float8_e5m2 if mode == "fp8_e5m2" else torch.float8_e4m3fn
    if t.dtype == qd:
        return t.to(torch.bfloat16)
    return t.to(torch.bfloat16).to(qd).to(torch.bfloat16)


def fp8_wraps(name, ndim):
    """Mirror of foley_fp8_wraps (csrc/safetensors.cuh): every >= 2-D "*.weight" of the DiT — the reference's deny
    list never matches because its recursion joins module names without dots (utils.py:441)."""
    return ndim >= 2 and name.endswith(".weight") and len(name) > len(".weight") and not name.startswith("dac.")
