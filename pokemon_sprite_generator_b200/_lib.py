"""ctypes binding of the C-ABI library (include/psg_b200.h).

There is no CPU fallback: if `libpsg_b200.so` is missing the import of any compute path raises.  The
structures below mirror csrc/gemm_epilogue.cuh and csrc/gemm_desc.h field for field.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

DT_F32, DT_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_SILU, ACT_MUL, ACT_RELU = 0, 1, 2, 3, 4
OP_KMAJOR, OP_MNMAJOR, OP_IM2COL, OP_IM2COL_T, OP_DGRAD, OP_CONVW_T = 0, 1, 2, 3, 4, 5

_TORCH_DT = {torch.float32: DT_F32, torch.bfloat16: DT_BF16}


class PsgEpilogue(C.Structure):
    _fields_ = [
        ("out", C.c_void_p), ("ldc", C.c_longlong), ("out_dtype", C.c_int), ("act_dtype", C.c_int),
        ("bias", C.c_void_p), ("rowbias", C.c_void_p), ("rows_per_group", C.c_int), ("ld_rowbias", C.c_longlong),
        ("act", C.c_int), ("alpha", C.c_float), ("residual", C.c_void_p), ("ldr", C.c_longlong),
        ("aux_out", C.c_void_p), ("aux_in", C.c_void_p), ("ld_aux", C.c_longlong), ("aux_act", C.c_int),
        ("accumulate", C.c_int), ("drop_seed", C.c_ulonglong), ("drop_threshold", C.c_uint), ("drop_scale", C.c_float),
    ]


class PsgOperand(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p), ("mode", C.c_int), ("ld", C.c_longlong),
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int), ("p", C.c_int), ("q", C.c_int),
        ("stride", C.c_int), ("pad", C.c_int), ("ksize", C.c_int), ("flip", C.c_int),
    ]


class PsgGemmDesc(C.Structure):
    _fields_ = [
        ("a", PsgOperand), ("b", PsgOperand), ("M", C.c_longlong), ("N", C.c_longlong), ("K", C.c_longlong),
        ("in_dtype", C.c_int), ("split_k", C.c_int), ("epi", PsgEpilogue),
    ]


class PsgError(RuntimeError):
    pass


_lib = None


def lib_file() -> Path:
    return Path(__file__).resolve().parent / "csrc" / "libpsg_b200.so"


def load() -> C.CDLL:
    """Load the library (building is the job of __graft_entry__.build / build.py, never done implicitly on a GPU box)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_file()
    if not path.exists():
        raise PsgError(
            f"{path} is missing: build it with `python -m pokemon_sprite_generator_b200.build` "
            "(there is no CPU or PyTorch fallback for the CUDA path)")
    lib = C.CDLL(str(path))
    lib.psg_last_error.restype = C.c_char_p
    lib.psg_version.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().psg_last_error().decode(errors="replace")
        raise PsgError(f"{what or 'psg call'} failed (rc={rc}): {msg}")


def raw_stream(device=None) -> int:
    """cudaStream_t of torch's current stream on `device` (default: the current device) as an integer.  The raw accessor, not
    torch.cuda.current_stream(): that builds a Stream object through several Python layers (device-index resolution, an
    os.getenv per call) and at ~1350 calls per train step was a fifth of the step's host time (tools/host_profile.py)."""
    idx = torch.cuda.current_device() if device is None else (device.index if device.index is not None else torch.cuda.current_device())
    return torch._C._cuda_getCurrentRawStream(idx)


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(raw_stream())


def ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def dt(t: torch.Tensor) -> int:
    return _TORCH_DT[t.dtype]


# When set to a list, every C call appends (name, start_event, end_event): a per-kernel-family time breakdown measured
# with CUDA events on the launching stream (tools/profile_step.py).
CALL_PROFILE = None


def call(name: str, *args) -> None:
    fn = getattr(load(), name)
    prof = CALL_PROFILE
    if prof is None:
        check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), name)
    e1.record()
    prof.append((name, e0, e1))
