"""ctypes binding of libmscl_b200.so -- the C-ABI declared in include/mscl_b200.h.

There is NO fallback: if the library is missing, fails to load, or the device is not
sm_100, every op raises.  Build it with `python -m mscl_b200.build` (nvcc, sm_100a).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# MSCL_LIB: another build of the same library (e.g. the MSCL_TIMELINE=1 debug build kept next to the product one)
LIB_PATH = os.environ.get("MSCL_LIB") or os.path.join(_HERE, "lib", "libmscl_b200.so")

c_int = ctypes.c_int32
c_i64 = ctypes.c_int64
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p

# name -> argtypes; every entry point returns int (0 = ok).  Mirrors include/mscl_b200.h.
PROTOTYPES = {
    "mscl_device_check": [c_int],
    "mscl_fetch_host": [c_ptr, c_ptr, c_i64, c_ptr],
    "mscl_enqueue": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr],
    "mscl_queue_export": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_ptr],
    "mscl_queue_import": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_ptr],
    "mscl_queue_weight": [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_i64, c_ptr],
    "mscl_ema_multi": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_f32, c_f32, c_ptr],
    "mscl_fra_maxrad": [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "mscl_fra_apply": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "mscl_fra_fused": [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "mscl_fra_rotate": [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_ptr],
    "mscl_hw_mean_fwd": [c_ptr, c_ptr, c_i64, c_int, c_ptr],
    "mscl_hw_mean_bwd": [c_ptr, c_ptr, c_i64, c_int, c_ptr],
    "mscl_hw_mean_ndhwc_fwd": [c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_ptr],
    "mscl_hw_mean_ndhwc_bwd": [c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_ptr],
    "mscl_lmcl": [c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "mscl_infonce_prep": [c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_i64, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_int, c_int,
                          c_ptr],
    "mscl_infonce_reduce_scatter": [c_ptr, c_int, c_int, c_int, c_ptr, c_int, c_ptr],
    "mscl_infonce_num_partials": [c_int, c_i64, c_int],
    "mscl_infonce_partial": [c_ptr, c_int, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_int, c_int, c_ptr],
    "mscl_infonce_partial_simt": [c_ptr, c_int, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_int, c_ptr],
    "mscl_infonce_reduce": [c_ptr, c_int, c_int, c_ptr, c_ptr],
    "mscl_infonce_finalize": [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_f32, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    "mscl_infonce_bwd": [c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr],
    "mscl_infonce_fused": [c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_i64, c_f32, c_f32, c_ptr, c_int, c_ptr, c_ptr, c_int, c_int,
                           c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr],
    "mscl_infonce_fused_multi": [c_int] + [c_ptr] * 13 + [c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "mscl_infonce_fused_multi_x": [c_int] + [c_ptr] * 13 + [c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_i64,
                                   c_int, c_int, c_ptr],
    "mscl_infonce_fused_parts_multi": [c_int, c_ptr, c_ptr, c_int],
    "mscl_infonce_bwd_slabs": [c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr],
    "mscl_infonce_bwd_slabs_multi": [c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "mscl_infonce_pass": [c_ptr, c_int, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_f32, c_ptr, c_int, c_int, c_int, c_ptr],
    "mscl_infonce_fused_parts": [c_int, c_i64, c_int],
    "mscl_gather_rows": [c_ptr, c_ptr, c_ptr, c_int, c_i64, c_ptr],
    "mscl_flow_visualize": [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr],
    "mscl_color_pipeline": [c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_int, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr],
    "mscl_grad_norm_multi": [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_f32, c_ptr, c_ptr, c_ptr],
    "mscl_clip_sgd_multi": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr, c_f32, c_f32, c_f32, c_int, c_ptr],
    "mscl_upsample_trilinear_fwd": [c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr],
    "mscl_upsample_trilinear_bwd": [c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr],
    "mscl_upsample_trilinear_ndhwc_fwd": [c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr],
    "mscl_upsample_trilinear_ndhwc_bwd": [c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr],
    "mscl_linear_axis_bwd": [c_ptr, c_ptr, c_i64, c_int, c_int, c_i64, c_ptr],
    "mscl_center_normalize": [c_ptr, c_i64, c_int, c_ptr, c_int, c_ptr, c_ptr, c_ptr],
    "mscl_retrieval_rank": [c_ptr, c_i64, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr],
}
EXPORTS = ["mscl_abi_version", "mscl_last_error"] + list(PROTOTYPES)

_lock = threading.Lock()
_lib = None
_launches = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)
# how many device kernels one successful call enqueues
_LAUNCHES_PER_CALL = {"mscl_device_check": 0, "mscl_infonce_num_partials": 0, "mscl_infonce_fused_parts": 0, "mscl_infonce_fused_parts_multi": 0, "mscl_color_pipeline": 2, "mscl_grad_norm_multi": 2,
                      "mscl_center_normalize": 3}


class MsclError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises MsclError with a build hint when absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MsclError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -m mscl_b200.build`.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.mscl_abi_version.restype = c_int
        lib.mscl_last_error.restype = ctypes.c_char_p
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = c_int
        if lib.mscl_abi_version() != 1:
            raise MsclError(f"ABI version mismatch: library reports {lib.mscl_abi_version()}")
        _lib = lib
    return _lib


# Optional device-side timing of selected entry points (bench.py's live roofline numbers):
# while active, every call to a selected entry point is bracketed by CUDA events on torch's
# current stream (the stream functional.py launches on) and filed with the algorithmic byte /
# flop count the caller states for that launch.
_timing = None


def start_timing(names):
    """Begin recording (start_event, end_event, algo_bytes, algo_flops) for the named entry points."""
    global _timing
    _timing = {n: [] for n in names}


def stop_timing():
    """Stop recording; returns {name: [(ms, algo_bytes, algo_flops), ...]}.  Synchronises the device."""
    global _timing
    import torch
    rec, _timing = _timing, None
    torch.cuda.synchronize()
    return {n: [(a.elapsed_time(b), nbytes, flops) for a, b, nbytes, flops in evs] for n, evs in (rec or {}).items()}


# NVTX: every entry-point call is a named range (nsys / ncu --nvtx timelines show the hot path's ops by name, the
# reference's tracing hook being mmcv's logger only).  MSCL_NVTX=0 turns the ranges off.
_nvtx = None


def _nvtx_api():
    global _nvtx
    if _nvtx is None:
        _nvtx = False
        if os.environ.get("MSCL_NVTX", "1") != "0":
            try:
                import torch
                if torch.cuda.is_available():
                    torch.cuda.nvtx.range_push("mscl_b200")
                    torch.cuda.nvtx.range_pop()
                    _nvtx = torch.cuda.nvtx
            except Exception:
                _nvtx = False
    return _nvtx


def call(name, *args, algo_bytes=0, algo_flops=0):
    """Invoke an entry point; raise MsclError carrying mscl_last_error() on failure.
    algo_bytes / algo_flops: algorithmic work of this launch, only filed
    when timing is active."""
    global _launches
    lib = load()
    timed = _timing is not None and name in _timing
    if timed:
        import torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    nvtx = _nvtx_api()
    if nvtx:
        nvtx.range_push(name)
    try:
        rc = getattr(lib, name)(*args)
    finally:
        if nvtx:
            nvtx.range_pop()
    if rc != 0:
        msg = lib.mscl_last_error()
        raise MsclError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")
    if timed:
        ev1.record()
        _timing[name].append((ev0, ev1, algo_bytes, algo_flops))
    _launches += _LAUNCHES_PER_CALL.get(name, 1)


def query(name, *args):
    """Invoke an entry point that returns a non-negative value (or a negative MSCL_E* code)."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc < 0:
        msg = lib.mscl_last_error()
        raise MsclError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")
    return rc


def launches():
    return _launches


_checked_devices = set()


def require_device(index):
    """Fail loudly unless CUDA device `index` is sm_100 (B200)."""
    if index in _checked_devices:
        return
    call("mscl_device_check", int(index))
    _checked_devices.add(index)
