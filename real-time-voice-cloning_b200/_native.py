"""ctypes binding of include/wavernn_b200.h.  Fails loudly when the CUDA library is missing or cannot be
loaded: the package has no CPU path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WRNN_B200_LIB") or os.path.join(HERE, "libwavernn_b200.so")   # override: A/B timing of kernel variants

LOOP_KERNELS = {0: "wrnn_loop_f32_kernel", 1: "wrnn_loop_tc_kernel", 2: "wrnn_loop_rs_kernel", 3: "wrnn_loop_sparse_kernel", 4: "wrnn_loop_tc2_kernel", 5: "wrnn_loop_rr_kernel", 6: "wrnn_loop_gn_kernel"}
OK, ERR_INVALID, ERR_NOT_LOADED, ERR_CUDA, ERR_TIMEOUT, ERR_SHAPE, ERR_TOO_SHORT = 0, -1, -2, -3, -4, -5, -6
MODE_RAW, MODE_MOL = 0, 1
TOPO_FATCHORD, TOPO_RUNTIMERACER, TOPO_GENEING = 0, 1, 2
PREC_F32, PREC_F16, PREC_SPARSE_F32 = 0, 1, 2
PREC_AUTO = -1        # host-side only: resolved per call by vocoder/models/fatchord_version.py:resolve_precision

PROGRESS_FN = C.CFUNCTYPE(None, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_void_p)


class Request(C.Structure):
    _fields_ = [
        ("n_utts", C.c_int32),
        ("mels", C.POINTER(C.c_void_p)),
        ("T", C.POINTER(C.c_int32)),
        ("mels_on_device", C.c_int32),
        ("batched", C.c_int32),
        ("target", C.c_int32),
        ("overlap", C.c_int32),
        ("mu_law", C.c_int32),
        ("apply_preemphasis", C.c_int32),
        ("precision", C.c_int32),
        ("seed", C.c_uint64),
        ("utt_index0", C.c_int32),
        ("fold_begin", C.c_int32),
        ("fold_end", C.c_int32),
        ("forced", C.c_void_p),
        ("max_steps", C.c_int32),
        ("progress", PROGRESS_FN),
        ("progress_user", C.c_void_p),
        ("wav", C.c_void_p),
        ("wav_capacity", C.c_int64),
        ("wav_offsets", C.POINTER(C.c_int64)),
        ("wav_on_device", C.c_int32),
        ("samples", C.c_void_p),
        ("logits", C.c_void_p),
        ("ms_h2d", C.c_float), ("ms_cond", C.c_float), ("ms_loop", C.c_float), ("ms_post", C.c_float), ("ms_d2h", C.c_float),
        ("n_folds", C.c_int32), ("n_steps", C.c_int32), ("n_launches", C.c_int32), ("loop_kernel", C.c_int32),
    ]


EXPORTS = ["wrnn_create", "wrnn_create_from_bin", "wrnn_set_topology", "wrnn_destroy", "wrnn_last_error", "wrnn_set_tensor", "wrnn_set_step", "wrnn_get_step",
           "wrnn_finalize", "wrnn_sparsity", "wrnn_sparse_available", "wrnn_fold_plan", "wrnn_generate", "wrnn_condition", "wrnn_condition_tc", "wrnn_postprocess",
           "wrnn_xfade_unfold", "wrnn_barrier_floor", "wrnn_cluster_floor", "wrnn_debug_umma_rate", "wrnn_debug_tc_gemm", "wrnn_debug_tc_gemm2", "wrnn_launch_count"]

_lib = None


def load():
    """Loads libwavernn_b200.so (built by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.wrnn_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.wrnn_set_topology.argtypes = [vp, C.c_int]
    lib.wrnn_create_from_bin.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp), C.c_char_p, C.c_int]
    lib.wrnn_destroy.argtypes = [vp]
    lib.wrnn_last_error.argtypes = [vp]
    lib.wrnn_last_error.restype = C.c_char_p
    lib.wrnn_set_tensor.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), C.c_int]
    lib.wrnn_set_step.argtypes = [vp, i64]
    lib.wrnn_get_step.argtypes = [vp]
    lib.wrnn_get_step.restype = i64
    lib.wrnn_finalize.argtypes = [vp]
    lib.wrnn_sparsity.argtypes = [vp]
    lib.wrnn_sparsity.restype = C.c_double
    lib.wrnn_sparse_available.argtypes = [vp]
    lib.wrnn_fold_plan.argtypes = [i64, i64, i64, C.POINTER(i64), C.POINTER(i64)]
    lib.wrnn_generate.argtypes = [vp, C.POINTER(Request)]
    lib.wrnn_condition.argtypes = [vp, vp, i32, vp, vp]
    lib.wrnn_condition_tc.argtypes = [vp, vp, i32, vp]
    lib.wrnn_postprocess.argtypes = [vp, vp, i64, i64, i32, i32, i32, i32, i32, vp]
    lib.wrnn_xfade_unfold.argtypes = [vp, vp, i64, i64, i32, vp]
    lib.wrnn_cluster_floor.argtypes = [vp, i32, i32, C.POINTER(C.c_float)]
    lib.wrnn_debug_umma_rate.argtypes = [vp, i32, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    lib.wrnn_debug_tc_gemm.argtypes = [vp, vp, vp, i32, vp]
    lib.wrnn_debug_tc_gemm2.argtypes = [vp, vp, vp, i32, vp]
    lib.wrnn_barrier_floor.argtypes = [vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.wrnn_launch_count.argtypes = [vp]
    lib.wrnn_launch_count.restype = i64
    _lib = lib
    return lib


def fold_plan(total_len, target, overlap):
    nf, padded = C.c_int64(), C.c_int64()
    rc = load().wrnn_fold_plan(total_len, target, overlap, C.byref(nf), C.byref(padded))
    if rc != OK:
        raise ValueError("bad fold plan arguments")
    return nf.value, padded.value
