"""ctypes binding of libfsw_embedding.so (the C ABI declared in include/fsw_embedding.h).

Mirrors how the reference binds its library (fsw_embedding.py:92-99, :196-206, :2952-2977): a
`ctypes.CDLL` next to the module, opened once per process.  Unlike the reference there is NO
pure-torch fallback: if the library is missing or fails to load, importing the product path raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSW_LIB_PATH") or os.path.join(_HERE, "libfsw_embedding.so")   # FSW_LIB_PATH: development builds (profiles/)

FSW_F32, FSW_F64 = 0, 1
PLAN_BUCKETS_PER_KIND = 519
PLAN_BUCKETS = 2 * PLAN_BUCKETS_PER_KIND

_lib = None

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int
c_vp = ctypes.c_void_p
c_sz = ctypes.c_size_t
c_dbl = ctypes.c_double

_SIGNATURES = {
    # name: (restype, argtypes)
    "fsw_version": (c_i32, []),
    "fsw_last_error": (ctypes.c_char_p, []),
    "fsw_built_for_sm": (c_i32, []),
    "fsw_launch_count": (c_i64, []),
    "fsw_profile_enable": (c_i32, [c_i32]),
    "fsw_profile_read": (c_i64, [ctypes.c_char_p, c_i64]),
    "segcumsum_wrapper": (None, [c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, ctypes.c_bool, c_i64, c_i64, c_sz]),
    "add_block_sums_wrapper": (None, [c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64]),
    "get_max_threads_per_block": (c_i32, [c_i32]),
    "launch_segcumsum_kernel_float": (None, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, ctypes.c_bool, c_i64, c_i64, c_i64]),
    "launch_segcumsum_kernel_double": (None, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, ctypes.c_bool, c_i64, c_i64, c_i64]),
    "launch_add_block_sums_kernel_float": (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64]),
    "launch_add_block_sums_kernel_double": (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64]),
    "fsw_segcumsum_workspace_bytes": (c_sz, [c_i64]),
    "fsw_segcumsum": (c_i32, [c_i32, c_vp, c_vp, c_vp, c_i32, c_i64, c_vp, c_sz, c_vp]),
    "fsw_csr_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "fsw_csr_from_edge_index": (c_i32, [c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "fsw_csr_coalesce_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "fsw_csr_coalesce": (c_i32, [c_i32, c_vp, c_i64, c_i64, c_i32, c_dbl, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "fsw_rowptr_from_sorted_rows": (c_i32, [c_vp, c_i64, c_i64, c_vp, c_vp]),
    "fsw_edge_weights": (c_i32, [c_i32, c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_dbl, c_i32, c_vp, c_vp, c_vp]),
    "fsw_plan_workspace_bytes": (c_sz, [c_i64]),
    "fsw_segment_plan": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_i64, c_dbl, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "fsw_gemm": (c_i32, [c_i32, c_i32, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i32, c_vp]),
    "fsw_gemm_fused": (c_i32, [c_i32, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_vp]),
    "fsw_set_tensor_cores": (c_i32, [c_i32]),
    "fsw_embed_scratch_bytes": (c_sz, [c_i32, c_vp, c_i64, c_i64, c_i32]),
    "fsw_embed_backward_extra_bytes": (c_sz, [c_i32, c_i64, c_i64]),
    "fsw_embed_forward": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64,
                                  c_vp, c_dbl, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_sz, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "fsw_embed_backward": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64,
                                   c_vp, c_dbl, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_sz, c_vp, c_i64,
                                   c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "fsw_embed_forward_cloud": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_dbl,
                                        c_vp, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp, c_vp, c_i64, c_vp]),
    "fsw_embed_backward_cloud": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp,
                                         c_vp, c_i64, c_vp]),
    "fsw_column_dot": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "fsw_embed_weight_grad_scratch_bytes": (c_sz, [c_i32, c_i64]),
    "fsw_embed_backward_weights": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_dbl, c_i32,
                                           c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp, c_sz, c_vp]),
    "fsw_embed_weight_grad_finish": (c_i32, [c_i32, c_vp, c_i64, c_vp, c_vp, c_i64, c_dbl, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "fsw_transpose_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "fsw_csr_transpose": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())


def load():
    """Open libfsw_embedding.so once per process.  Raises RuntimeError when it cannot be loaded
    (the reference's behaviour with fsw_embedding_produce_error_on_custom_library_loading_failure,
    fsw_embedding.py:201-203); there is deliberately no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libfsw_embedding.so not found at %s - build it with `python -m fsw_gnn_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU / torch fallback" % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise RuntimeError("Error loading CUDA library '%s': %s" % (LIB_PATH, e))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().fsw_last_error().decode("utf-8", "replace")
        raise RuntimeError("libfsw_embedding %s failed (code %d): %s" % (what, rc, msg))


def dtype_code(dtype):
    if dtype == torch.float32:
        return FSW_F32
    if dtype == torch.float64:
        return FSW_F64
    raise TypeError("libfsw_embedding supports float32 and float64, got %s" % dtype)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def call(device, name, *args):
    """Run library entry point `name` with `device` as the current CUDA device and raise on a non-zero status.
    The C ABI takes raw pointers and a stream; launches, memsets and function-attribute calls act on the calling
    thread's CURRENT device, so every entry into the library runs in the device context of the tensors it is given
    (a module on cuda:1 works whatever the current device is, like the torch reference)."""
    fn = getattr(load(), name)
    with torch.cuda.device(device):
        check(fn(*args), name)


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the FSW kernels run on the GPU only (no CPU fallback)" % name)


def launch_count():
    return int(load().fsw_launch_count())


def profile_enable(on=True):
    load().fsw_profile_enable(1 if on else 0)


def profile_read():
    """{label: (count, total_ms)} of the launches since the last read (synchronises their events)."""
    lib = load()
    cap = 1 << 16  # one call: reading consumes the records
    buf = ctypes.create_string_buffer(cap)
    lib.fsw_profile_read(buf, cap)
    out = {}
    for line in buf.value.decode().splitlines():
        label, cnt, ms = line.split()
        out[label] = (int(cnt), float(ms))
    return out
