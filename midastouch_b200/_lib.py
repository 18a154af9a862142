"""ctypes binding of libmidas_b200.so (include/midas_b200.h).

There is no fallback: if the shared library is missing or a call fails this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MIDAS_B200_LIB") or os.path.join(_HERE, "libmidas_b200.so")  # override: A/B builds only
SOURCES = [os.path.join(_HERE, "csrc", "midas_b200.cu")]
HEADERS = [os.path.join(_HERE, "csrc", "mt_math.cuh"), os.path.join(_HERE, "csrc", "mt_nn.cuh"), os.path.join(_HERE, "csrc", "mt_mesh.cuh"), os.path.join(_HERE, "csrc", "mt_tcn.cuh"), os.path.join(_HERE, "csrc", "mt_cluster.cuh"), os.path.join(_HERE, "csrc", "mt_dbscan.cuh"), os.path.join(_HERE, "csrc", "mt_gemm_tma.cuh"), os.path.join(_HERE, "..", "include", "midas_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


class MidasError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    deps = SOURCES + HEADERS
    if (not force) and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    subprocess.check_call(cmd)
    return LIB_PATH


class StepArgs(C.Structure):
    _fields_ = [
        ("d_soa_cur", C.c_void_p), ("d_soa_next", C.c_void_p), ("stride", C.c_longlong),
        ("d_nn_cur", C.c_void_p), ("d_nn_next", C.c_void_p), ("d_anc", C.c_void_p), ("n", C.c_longlong),
        ("odom", C.c_float * 16), ("d_tn", C.c_void_p), ("d_rot", C.c_void_p),
        ("sig_t", C.c_float), ("sig_r", C.c_float),
        ("seed", C.c_uint64), ("step", C.c_uint64), ("first_gid", C.c_uint64),
        ("softmax", C.c_int), ("u", C.c_float), ("resample", C.c_int),
        ("gt", C.c_void_p), ("d_rmse2", C.c_void_p),
        ("rank", C.c_int), ("world", C.c_int), ("n_global", C.c_longlong),
        ("d_shard_sums", C.c_void_p), ("d_n_out", C.c_void_p), ("d_n_in", C.c_void_p),
        ("prune_dist", C.c_double), ("d_cb_poses", C.c_void_p), ("table_ready_event", C.c_void_p), ("fuse_sums", C.c_int),
    ]


_SIGS = {
    "mt_last_error": (C.c_char_p, []),
    "mt_version": (C.c_int, []),
    "mt_ctx_create": (C.c_int, [C.c_int, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mt_ctx_destroy": (C.c_int, [C.c_void_p]),
    "mt_codebook_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mt_codebook_grid_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mt_codebook_nbr_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "mt_codebook_rank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mt_ctx_set_timing_events": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mt_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mt_step_graph_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    "mt_dbscan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_longlong, C.c_void_p, C.POINTER(C.c_int), C.c_void_p]),
    "mt_ctx_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]),
    "mt_mesh_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double]),
    "mt_prune_aos": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mt_codebook_query": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mt_codebook_query_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mt_cosine_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "mt_cosine_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "mt_softmax_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "mt_aos_to_soa": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p]),
    "mt_soa_to_aos": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]),
    "mt_se3_keys": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]),
    "mt_se3_keys_w": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_double, C.c_void_p, C.c_void_p]),
    "mt_nn_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "mt_nn_assign": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mt_gather_rows_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "mt_motion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_float, C.c_float, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]),
    "mt_rmse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mt_resample_systematic": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mt_resample_multinomial": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_uint64, C.c_uint64, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "mt_gather_soa": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p]),
    "mt_gather_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "mt_cluster_centers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mt_select_k": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mt_tcn_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mt_tcn_destroy": (C.c_int, [C.c_void_p]),
    "mt_tcn_set_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mt_tcn_set_bn": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float]),
    "mt_tcn_set_gem": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "mt_tcn_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mt_tcn_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mt_dist_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mt_dist_import": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mt_dist_debug": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong)]),
    "mt_trace_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]),
    "mt_step_is_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "mt_step_a": (C.c_int, [C.c_void_p, C.POINTER(StepArgs), C.c_void_p]),
    "mt_step_local_sum_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mt_step_b": (C.c_int, [C.c_void_p, C.POINTER(StepArgs), C.c_void_p]),
    "mt_step_weights": (C.c_int, [C.c_void_p, C.POINTER(StepArgs), C.c_void_p, C.c_void_p]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MidasError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  midastouch_b200 has no CPU or PyTorch fallback."
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise MidasError(f"{what} failed ({rc}): {lib().mt_last_error().decode()}")


def call(name: str, *args):
    check(getattr(lib(), name)(*args), name)


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    """device/host pointer of a contiguous tensor (0 for None)."""
    if t is None:
        return 0
    if not t.is_contiguous():
        raise MidasError("tensor passed to libmidas_b200 must be contiguous")
    return t.data_ptr()
