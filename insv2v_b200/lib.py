"""ctypes binding of libivv_b200.so (include/ivv.h). The product path has no fallback: if the library is missing or a
call fails, a RuntimeError is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IVV_LIB_PATH") or os.path.join(_HERE, "libivv_b200.so")  # override: tuning A/B only

c_void_p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_f32 = ctypes.c_float
c_size = ctypes.c_size_t


class GemmArgs(ctypes.Structure):
    """Mirror of ivv_gemm_args (include/ivv.h)."""
    _fields_ = [
        ("a", c_void_p), ("n_img", c_i64), ("h", c_i64), ("w", c_i64), ("c", c_i64), ("a_ld", c_i64),
        ("wgt", c_void_p), ("n_out", c_i64), ("w_ld", c_i64),
        ("taps", c_i32), ("geglu", c_i32),
        ("d", c_void_p), ("d_ld", c_i64), ("out_f32", c_i32), ("splits", c_i32),
        ("bias", c_void_p), ("rowbias", c_void_p), ("rowbias_group", c_i64), ("rowbias_ld", c_i64),
        ("residual", c_void_p), ("res_ld", c_i64),
        ("tap_h", c_i32), ("tap_w", c_i32), ("relu", c_i32), ("rowbias_mod", c_i32),
        ("row_stats_out", c_void_p), ("ln_stats", c_void_p), ("ln_wsum", c_void_p), ("ln_parts", c_i32), ("ln_eps", c_f32),
    ]


# name -> (restype, argtypes); every symbol include/ivv.h declares
SIGNATURES = {
    "ivv_abi_version": (c_i32, []),
    "ivv_last_error": (ctypes.c_char_p, []),
    "ivv_gemm": (c_i32, [ctypes.POINTER(GemmArgs), c_void_p]),
    "ivv_gemm_ln_fold_ok": (c_i32, [c_i64, c_i64, c_i64]),
    "ivv_splitk_reduce": (c_i32, [c_void_p, c_i32, c_i64, c_i64, c_i64, c_void_p, c_void_p, c_i64, c_i64, c_void_p,
                                  c_i64, c_void_p, c_i64, c_void_p]),
    "ivv_im2col_s2": (c_i32, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_i32, c_void_p]),
    "ivv_groupnorm": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i32, c_i64, c_f32, c_i32,
                              c_void_p, c_size, c_void_p]),
    "ivv_groupnorm_ws_bytes": (c_size, [c_i64, c_i32, c_i64]),
    "ivv_groupnorm_is_fused": (c_i32, [c_i64, c_i64, c_i64, c_i32, c_i64]),
    "ivv_groupnorm2": (c_i32, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i32, c_i64,
                               c_f32, c_i32, c_void_p, c_size, c_void_p]),
    "ivv_layernorm": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_f32, c_void_p, c_i64, c_i64,
                              c_i64, c_void_p]),
    "ivv_attention": (c_i32, [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64,
                              c_i32, c_i32, c_f32, c_void_p]),
    "ivv_temporal_attention": (c_i32, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i32, c_f32, c_void_p]),
    "ivv_softmax_rows": (c_i32, [c_void_p, c_i32, c_void_p, c_i64, c_i64, c_f32, c_void_p]),
    "ivv_upsample_nearest": (c_i32, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "ivv_concat_channels": (c_i32, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p]),
    "ivv_ncfhw_to_frames": (c_i32, [c_void_p, c_i32, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "ivv_frames_to_ncfhw": (c_i32, [c_void_p, c_i32, c_i64, c_void_p, c_i32, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "ivv_timestep_embedding": (c_i32, [c_void_p, c_void_p, c_i64, c_i32, c_i32, c_f32, c_void_p]),
    "ivv_silu": (c_i32, [c_void_p, c_void_p, c_i64, c_void_p]),
    "ivv_scale": (c_i32, [c_void_p, c_void_p, c_i64, c_f32, c_f32, c_void_p]),
    "ivv_warp_image": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "ivv_resize_flow": (c_i32, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "ivv_flow_noise_correction": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "ivv_cfg_ddim_step": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_f32, c_f32, c_f32, c_f32, c_void_p]),
    "ivv_frames_u8_to_f32": (c_i32, [c_void_p, c_void_p, c_i64, c_i64, c_i32, c_void_p]),
    "ivv_frames_to_u8": (c_i32, [c_void_p, c_i32, c_void_p, c_i64, c_i64, c_void_p]),
    "ivv_sampler_begin": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64,
                                  c_void_p]),
    "ivv_sampler_partials": (c_i64, [c_i64, c_i64]),
    "ivv_sampler_combine": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_i64, c_i64,
                                    c_void_p]),
    "ivv_sampler_update": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    # RAFT optical flow
    "ivv_im2col": (c_i32, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i64,
                           c_i64, c_void_p]),
    "ivv_channelnorm": (c_i32, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_f32, c_i32,
                                c_void_p, c_void_p, c_size, c_void_p]),
    "ivv_channelnorm_ws_bytes": (c_size, [c_i64, c_i64, c_i64]),
    "ivv_add_relu": (c_i32, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "ivv_raft_prep_images": (c_i32, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_void_p]),
    "ivv_avgpool2_f32": (c_i32, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p]),
    "ivv_corr_lookup": (c_i32, [ctypes.POINTER(c_void_p), c_i32, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i32,
                                c_f32, c_void_p]),
    "ivv_raft_init_state": (c_i32, [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_i64, c_i32, c_i32, c_void_p]),
    "ivv_gru_gate_r": (c_i32, [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_i32, c_void_p]),
    "ivv_gru_update": (c_i32, [c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i32, c_void_p]),
    "ivv_raft_update_coords": (c_i32, [c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64,
                                       c_void_p]),
    "ivv_convex_upsample": (c_i32, [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p]),
}

ABI_VERSION = 6  # IVV_ABI_VERSION of include/ivv.h
_lib = None
LAUNCH_COUNT = 0  # incremented by ops.py for every kernel-launching C-ABI call (bench.py reports it)


def load():
    """Load the shared library (building is __graft_entry__.build()'s job). Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.ivv_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libivv_b200.so ABI version {lib.ivv_abi_version()} != {ABI_VERSION} (stale build?)")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().ivv_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")
