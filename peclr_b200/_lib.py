"""ctypes binding of libpeclr_b200.so (the C ABI in include/peclr_b200.h).

There is no fallback: if the library is missing or a kernel returns an error the call raises.
"""
import ctypes
import os
from ctypes import c_double, c_float, c_int, c_longlong, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# PECLR_B200_LIB: load another build of the same ABI (A/B timing of kernel changes); never a fallback
LIB_PATH = os.environ.get("PECLR_B200_LIB") or os.path.join(_HERE, "libpeclr_b200.so")

ABI_VERSION = 3  # include/peclr_b200.h

P = c_void_p
I = c_int
L = c_longlong
F = c_float

# name -> argtypes (return type is int unless listed in _RESTYPES)
SIGNATURES = {
    "peclr_abi_version": [],
    "peclr_conv2d_fprop": [P, P, P, I, I, I, I, I, I, I, P, P, P],
    "peclr_conv2d_dgrad": [P, P, P, I, I, I, I, I, I, I, I, P],
    "peclr_conv2d_dgrad_bnreduce": [P, P, P, I, I, I, I, I, I, I, P, P, P, P, P, P, P],
    "peclr_conv2d_dgrad_finish": [P, P, P, I, I, I, I, I, P, P, P, P],
    "peclr_conv2d_dgrad_finish_lattice": [P, P, P, I, I, I, I, I, P, P, P, I, P],
    "peclr_conv2d_wgrad_workspace_bytes": [I, I, I, I, I, I, I],
    "peclr_conv2d_wgrad": [P, P, P, I, I, I, I, I, I, I, P, L, P],
    "peclr_conv2d_wgrad_splits": [I, I, I, I, I, I, I],
    "peclr_conv2d_wgrad_partials": [P, P, P, I, I, I, I, I, I, I, P, L, P],
    "peclr_wgrad_reduce_block_f4": [],
    "peclr_wgrad_reduce_batched": [P, I, I, P],
    "peclr_stem_fprop": [P, P, P, I, I, I, P, P, P],
    "peclr_stem_wgrad_workspace_bytes": [I, I, I],
    "peclr_stem_wgrad": [P, P, P, I, I, I, P, L, P],
    "peclr_bn_apply": [P] * 20 + [L, I, F, F, I, P],
    "peclr_bn_bwd_reduce": [P, P, P, P, P, P, P, I, P, L, I, P],
    "peclr_bn_bwd_apply": [P, P, P, P, P, P, P, I, P, P, P, P, P, L, I, P],
    "peclr_stem_bn_relu_pool": [P] * 11 + [I, I, I, F, F, P],
    "peclr_stem_pool_bwd": [P] * 9 + [I, I, I, P],
    "peclr_avgpool_fwd": [P, P, I, I, I, P],
    "peclr_avgpool_bwd": [P, P, I, I, I, P],
    "peclr_stem_input": [P, P, P, I, I, I, P],
    "peclr_sgemm_workspace_bytes": [I, I, I],
    "peclr_sgemm": [P, P, P, P, I, I, I, L, L, L, L, L, I, P, L, P],
    "peclr_bn1d_relu_fwd": [P] * 8 + [I, I, F, F, P],
    "peclr_bn1d_relu_bwd": [P] * 9 + [I, I, P],
    "peclr_colsum_acc": [P, P, I, I, P],
    "peclr_ntxent_workspace_bytes": [I, I],
    "peclr_ntxent_fused": [P, P, P, P, I, I, I, I, I, I, F, P, P, P, P, L, I, I, P, P, P],
    "peclr_ntxent_plain": [P, I, I, F, P, P, P, L, P],
    "peclr_translate_encodings": [P, P, P, I, I, I, I, P],
    "peclr_rotate_encoding": [P, P, P, I, I, I, P],
    "peclr_rotate_encoding_bwd": [P, P, I, I, I, P],
    "peclr_rotation_2d_matrix": [P, P, P, c_double, P, I, P],
    "peclr_projection_stats": [P, P, I, I, I, P],
    "peclr_opt_chunk_elems": [],
    "peclr_lars_adam_step": [P, P, P, P, P, P, P, I, P, P, I, P, F, I, F, F, F, I, F, I, F, P],
    "peclr_cast_bf16": [P, P, L, P],
    "peclr_weight_transpose": [P, P, P, I, I, P],
    "peclr_stem_pack": [P, P, P],
    "peclr_stem_unpack_grad": [P, P, P],
    "peclr_rn25d_head": [P, P, I, I, P, F, F, P, P, P, P, P],
    "peclr_two_view_augment": [P, L, P, I, I, I, F, F, F, F, F, F, P, P, P],
}
_RESTYPES = {"peclr_ntxent_workspace_bytes": c_longlong, "peclr_conv2d_wgrad_workspace_bytes": c_longlong,
             "peclr_stem_wgrad_workspace_bytes": c_longlong, "peclr_sgemm_workspace_bytes": c_longlong}
_NO_CHECK = {"peclr_abi_version", "peclr_ntxent_workspace_bytes", "peclr_opt_chunk_elems", "peclr_conv2d_wgrad_splits",
             "peclr_wgrad_reduce_block_f4",
             "peclr_conv2d_wgrad_workspace_bytes", "peclr_stem_wgrad_workspace_bytes", "peclr_sgemm_workspace_bytes"}

_ERRORS = {-1001: "bad argument", -1002: "CUDA driver entry point unavailable", -1003: "TMA tensor-map encoding failed"}


class PeclrKernelError(RuntimeError):
    pass


_lib = None


def load():
    """Loads the shared library (raises if it has not been built: run `python -m peclr_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PeclrKernelError(
            f"{LIB_PATH} is missing: the CUDA extension must be built (python -m peclr_b200.build); "
            "peclr_b200 has no CPU / PyTorch fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    if lib.peclr_abi_version() != ABI_VERSION:
        raise PeclrKernelError(f"{LIB_PATH} implements ABI version {lib.peclr_abi_version()}, this binding needs "
                               f"{ABI_VERSION}: rebuild it (python -m peclr_b200.build --force)")
    _lib = lib
    return lib


def _check(name, rc):
    if rc != 0:
        what = _ERRORS.get(rc, f"cudaError {-rc}" if rc < 0 else "unknown")
        raise PeclrKernelError(f"{name} failed: {rc} ({what})")


LAUNCHES = 0  # kernels of this library launched so far (bench.py reports the count of the timed region)
_PROFILE = None


def _kernels_in_call(name, args):
    if name in ("peclr_conv2d_dgrad", "peclr_conv2d_dgrad_bnreduce"):
        return 4 if (args[8] == 3 and args[9] == 2) else 1
    if name == "peclr_lars_adam_step":
        return 2 if args[17] else 1
    if name == "peclr_conv2d_wgrad_partials":
        return 1
    if name == "peclr_conv2d_wgrad":  # + the ordered reduction of the pixel splits
        return 2 if load().peclr_conv2d_wgrad_workspace_bytes(*args[3:10]) > 0 else 1
    if name == "peclr_stem_wgrad":
        return 2 if load().peclr_stem_wgrad_workspace_bytes(*args[3:6]) > 0 else 1
    return 1


def call(name, *args):
    """Calls an entry point, converting torch tensors to device pointers; raises on a non-zero return."""
    global LAUNCHES
    lib = load()
    conv = []
    for a in args:
        if hasattr(a, "data_ptr"):
            conv.append(a.data_ptr())
        else:
            conv.append(a)
    if name in _NO_CHECK:
        return getattr(lib, name)(*conv)
    prof = _PROFILE is not None and name in _PROFILE["names"]
    if prof:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # keep the GPU busy (~100 us spin) while the host prepares the launch, so that the event pair brackets
        # the kernel only and not the host-side launch latency
        torch.cuda._sleep(200000)
        e0.record()
    rc = getattr(lib, name)(*conv)
    if prof:
        e1.record()
        _PROFILE["records"].append((name, args, e0, e1))
    _check(name, rc)
    LAUNCHES += _kernels_in_call(name, args)
    return rc


def profile_calls(fn, names):
    """Runs fn() with every call to the named entry points bracketed by CUDA events on the launching stream.
    Returns [(name, args, milliseconds)]."""
    global _PROFILE
    import torch

    _PROFILE = {"names": set(names), "records": []}
    try:
        fn()
        torch.cuda.synchronize()
        return [(n, a, e0.elapsed_time(e1)) for n, a, e0, e1 in _PROFILE["records"]]
    finally:
        _PROFILE = None


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream
