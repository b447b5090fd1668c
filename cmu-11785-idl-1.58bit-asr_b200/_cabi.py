"""ctypes binding of libonebit.so (include/onebit.h).  Plain pointers and sizes only.

The library is built in-tree by ``csrc/build.sh`` (``__graft_entry__.build()``).  There is no
fallback of any kind: if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libonebit.so")

OB_OK, OB_ERR_ARG, OB_ERR_CUDA, OB_ERR_ARCH, OB_ERR_WORKSPACE = 0, 1, 2, 3, 4
OB_F32, OB_BF16 = 0, 1
OB_ALPHA_EFF, OB_ALPHA_RAW = 0, 1
OB_ORDER_I8, OB_ORDER_BF16 = 0, 1
OB_PREP_TAIL, OB_PREP_SWISH = 1, 2
DBG_SWAP_LBO_SBO, DBG_FORCE_BLOCK_N, DBG_FORCE_SPLITS, DBG_MAX_CTAS = 1, 2, 3, 4
DBG_SMALL_M = 9         # 1: keep M <= 64 on the tcgen05 kernel instead of the weight-streaming DP4A kernel

_p, _i, _i64, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
_u64, _u32 = ctypes.c_uint64, ctypes.c_uint32

# name -> (restype, argtypes); mirrors include/onebit.h one to one
SIGNATURES = {
    "ob_version": (_i, []),
    "ob_launch_count": (_i64, []),
    "ob_last_error_string": (ctypes.c_char_p, []),
    "ob_absmean_workspace_bytes": (_sz, []),
    "ob_weight_absmean": (_i, [_p, _i64, _p, _p, _p]),
    "ob_weight_quant_pack": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "ob_weight_quant_pack_multi": (_i, [_p, _i, _i, _i, _p]),
    "ob_weight_quant_dense": (_i, [_p, _p, _i, _i64, _i, _p, _p]),
    "ob_ste_workspace_bytes": (_sz, [_i64]),
    "ob_weight_ste_backward": (_i, [_p, _p, _p, _i, _i64, _i, _p, _p, _p, _p]),
    "ob_unpack_codes": (_i, [_p, _i, _i, _i, _p, _p]),
    "ob_act_quant_i8": (_i, [_p, _i, _i64, _i, _p, _p, _p]),
    "ob_gemm_tern_i8_fwd": (_i, [_p, _p, _p, _p, _i, _p, _i, _i, _i, _p, _i, _p]),
    "ob_bwd_colsum_blocks": (_i, [_i]),
    "ob_bwd_prep": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "ob_bwd_dx": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _p]),
    "ob_bwd_dw_workspace_bytes": (_sz, [_i, _i, _i]),
    "ob_bwd_dw": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "ob_bwd_dw_q8": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "ob_bwd_dw_q8_groups": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "ob_swish_drop_quant": (_i, [_p, _p, ctypes.c_float, _u64, _u64, _u32, _i64, _i, _p, _p, _p]),
    "ob_swish_drop_bwd": (_i, [_p, _p, _p, ctypes.c_float, _u64, _u64, _u32, _i64, _p, _p]),
    "ob_layernorm_fwd": (_i, [_p, _p, _p, ctypes.c_float, _i64, _i, _p, _p, _p, _p]),
    "ob_layernorm_bwd_workspace_bytes": (_sz, [_i]),
    "ob_layernorm_bwd": (_i, [_p, _p, _p, _p, _p, _i64, _i, _p, _p, _p, _p, _p]),
    "ob_layernorm_quant_fwd": (_i, [_p, _p, _p, ctypes.c_float, _i64, _i, _p, _p, _p, _p, _p]),
    "ob_layernorm_bwd3": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _p, _p, _p, _p, _p]),
    "ob_gemm_tern_i8_fwd_tail": (_i, [_p, _p, _p, _p, _i, _p, _i, _i, _i, _p, _p, ctypes.c_float, _u64, _u64, _u32, _i64, _p,
                                      _p]),
    "ob_bwd_prep_fused": (_i, [_p, _i, _p, _p, ctypes.c_float, _u64, _u64, _u32, _i64, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "ob_relattn_softmax_fwd": (_i, [_p, _p, _p, _p, ctypes.c_float, _u64, _u64, _u32, ctypes.c_float, _i, _i, _i, _i, _p, _p,
                                    _p]),
    "ob_relattn_softmax_bwd": (_i, [_p, _p, _p, ctypes.c_float, _u64, _u64, _u32, ctypes.c_float, _i, _i, _i, _i, _p, _p,
                                    _p]),
    "ob_ctc_decode_workspace_bytes": (_sz, [_i, _i]),
    "ob_ctc_greedy_decode": (_i, [_p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _p]),
    "ob_gemm_f32_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "ob_gemm_f32": (_i, [_p, _i, _i64, _i64, _i64, _p, _i, _i64, _i64, _i64, _p, _i64, _i64, _i64, _p, ctypes.c_float, _i,
                         _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "ob_convmod_workspace_bytes": (_sz, [_i, _i, _i]),
    "ob_glu_dwconv_bn_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, ctypes.c_float, _i, _p, _p, _p, _p, _p]),
    "ob_bn_swish_fwd": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _p, _p]),
    "ob_bn_swish_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _i, _p, _p, _p, _p]),
    "ob_glu_dwconv_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ob_add_bias2": (_i, [_p, _p, _p, _i64, _i, _p, _p, _p]),
    "ob_add_colsum2_workspace_bytes": (_sz, [_i64, _i]),
    "ob_add_colsum2": (_i, [_p, _p, _i64, _i, _p, _p, _p, _p]),
    "ob_residual_dropout_fwd": (_i, [_p, _p, _p, ctypes.c_float, ctypes.c_float, _u64, _u64, _u32, _i64, _i, _p, _p]),
    "ob_residual_dropout_bwd": (_i, [_p, _p, ctypes.c_float, ctypes.c_float, _u64, _u64, _u32, _i64, _i, _p, _p]),
    "ob_conv1_relu_workspace_bytes": (_sz, []),
    "ob_conv1_relu_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "ob_conv1_relu_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ob_ctc_state_pitch": (_i, [_i]),
    "ob_ctc_loss_fwd": (_i, [_p, _i64, _p, _p, _i64, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "ob_ctc_loss_bwd": (_i, [_p, _i64, _p, _p, _i64, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ob_colsum_workspace_bytes": (_sz, [_i64, _i]),
    "ob_colsum": (_i, [_p, _i64, _i, _p, _p, _p]),
    "ob_debug_set": (_i, [_i, _i]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()' or csrc/build.sh). "
            "There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()
if os.environ.get("OB_PDL", "1") == "0":       # A/B switch: launch every kernel fully serialised (no programmatic dependent launch)
    lib.ob_debug_set(13, 0)


def last_error() -> str:
    return lib.ob_last_error_string().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a library return code to a Python exception (bad arguments -> ValueError)."""
    if rc == OB_OK:
        return
    msg = last_error()
    if rc == OB_ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError(f"libonebit error {rc}: {msg}")
