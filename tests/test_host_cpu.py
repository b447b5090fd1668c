"""CPU-only checks: the C-ABI library loads and exports every symbol include/onebit.h declares, argument
validation happens before any device work, and the host-side mirror keeps the reference's interface."""
import ctypes
import hashlib
import math
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, bits_to_f32

import onebit_b200 as ob
from onebit_b200 import _cabi


def header_symbols():
    text = open(os.path.join(ROOT, "include", "onebit.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ob_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 17
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/onebit.h but not exported"
    assert set(syms) == set(_cabi.SIGNATURES), "ctypes table and header disagree"
    assert _cabi.lib.ob_version() == 100


def test_argument_validation_without_gpu():
    lib = _cabi.lib
    # bad bitwidth -> OB_ERR_ARG with the reference's message (quant.py:66)
    rc = lib.ob_weight_quant_pack(16, 16, 1, 64, 64, 3, 16, None, None)
    assert rc == _cabi.OB_ERR_ARG and "bitwidth must be one of {1,2,32}" in _cabi.last_error()
    with pytest.raises(ValueError):
        _cabi.check(rc)
    assert lib.ob_act_quant_i8(None, 0, 4, 64, None, None, None) == _cabi.OB_ERR_ARG
    assert lib.ob_act_quant_i8(16, 0, 4, 70, 16, 16, None) == _cabi.OB_ERR_ARG          # K % 16
    assert lib.ob_gemm_tern_i8_fwd(16, 16, 16, 16, 1, None, 8, 64, 100, 16, 0, None) == _cabi.OB_ERR_ARG  # K % 64
    assert lib.ob_bwd_dx(16, 16, 16, 16, 1, 8, 60, 64, 16, 0, None) == _cabi.OB_ERR_ARG                   # N % 64
    assert lib.ob_debug_set(999, 1) == _cabi.OB_ERR_ARG
    assert lib.ob_bwd_dw_workspace_bytes(1000, 256, 256) >= 256 * 256 * 4
    assert lib.ob_bwd_colsum_blocks(129) == (129 + 31) // 32
    # grad_W from the int8 codes: argument checks come before the device is touched
    dw8 = lambda **kw: lib.ob_bwd_dw_q8_groups(16, 16, None, 16, 16, 1, kw.get("rows2", 10), kw.get("M", 100), kw.get("N", 256),  # noqa: E731
                                               kw.get("K", 256), 16, 16, kw.get("gb", None), kw.get("ws", 16), kw.get("ws_bytes", 1 << 30), None)
    assert dw8(rows2=101) == _cabi.OB_ERR_ARG and "rows2" in _cabi.last_error()
    assert dw8(rows2=-1) == _cabi.OB_ERR_ARG and dw8(K=100) == _cabi.OB_ERR_ARG and dw8(ws=None) == _cabi.OB_ERR_ARG
    assert dw8(gb=16) == _cabi.OB_ERR_ARG and "colsum" in _cabi.last_error()                     # grad_bias without column sums
    assert dw8(ws_bytes=64) == _cabi.OB_ERR_WORKSPACE
    assert lib.ob_bwd_dw_q8(16, 16, None, 16, 16, 1, 3, 100, 256, 256, 16, 16, None, 16, 1 << 30, None) == _cabi.OB_ERR_ARG   # bitwidth
    # entry points around the layer: shapes and strides are checked before anything touches the device
    E = _cabi.OB_ERR_ARG
    gemm = lambda **kw: lib.ob_gemm_f32(16, 0, kw.get("lda", 64), 0, 0, 32, 0, 64, 0, 0, 48, kw.get("ldd", 64), 0, 0, None, 1.0, 0,  # noqa: E731
                                        8, 8, kw.get("K", 64), 1, 1, kw.get("passes", 3), None, 0, None)
    assert gemm(passes=2) == E and "passes" in _cabi.last_error()
    assert gemm(lda=62) == E and gemm(K=0) == E and gemm(ldd=4) == E
    assert lib.ob_gemm_f32_workspace_bytes(5004, 256, 25536, 1, 1) == 25 * 5004 * 256 * 4     # 25 K-chunks of 1024
    assert lib.ob_gemm_f32_workspace_bytes(25536, 5004, 256, 1, 1) == 0 and lib.ob_gemm_f32_workspace_bytes(399, 399, 64, 64, 4) == 0
    assert lib.ob_glu_dwconv_bn_fwd(16, 16, None, 6, 50, 64, 33, 1e-5, 1, 16, 16, 16, 16, None) == E         # 33 taps
    assert lib.ob_glu_dwconv_bn_fwd(16, 16, None, 6, 50, 64, 31, 1e-5, 4, 16, 16, 16, 16, None) == E         # 6 % 4 groups
    assert lib.ob_bn_swish_fwd(16, 16, 16, 16, 16, 100, 60, 1, 16, None) == E                                # C % 64
    assert lib.ob_residual_dropout_fwd(16, 16, None, 1.0, 1.0, 0, 0, 70000, 8, 64, 16, None) == E             # 16-bit threshold
    assert lib.ob_swish_drop_quant(16, None, 1.0, 0, 0, 0, 8, 300, 16, 16, None) == E                        # K not supported
    assert lib.ob_relattn_softmax_fwd(16, 16, 16, None, 1.0, 0, 0, 0, 0.125, 1, 1, 40, 39, 16, None, None) == E   # ld < T
    assert lib.ob_conv1_relu_fwd(16, 16, None, 2, 50, 80, 128, 16, None) == E                                 # C != 256
    assert lib.ob_convmod_workspace_bytes(64, 399, 256) >= 64 * 7 * 32 * 256 * 4
    assert lib.ob_ctc_state_pitch(64) == 132 and lib.ob_ctc_state_pitch(0) == 4
    ctc_fwd = lambda **kw: lib.ob_ctc_loss_fwd(16, kw.get("ld", 8), 16, 16, 3, 16, 2, 4, 8, kw.get("L", 3), kw.get("blank", 0),  # noqa: E731
                                               16, 16, 16, 16, 16, None)
    assert ctc_fwd(blank=8) == E and ctc_fwd(ld=7) == E and ctc_fwd(L=512) == E                              # blank >= V, pitch < V, L > 511


def test_constructor_matches_reference_fixtures(kat_seeded):
    for key, ref in kat_seeded.items():
        torch.manual_seed(0)
        m = ob.QuantizedLinear(ref["in"], ref["out"])
        assert hashlib.sha256(m.weight.detach().numpy().tobytes()).hexdigest()[:16] == ref["sha_W"], key
        assert math.isclose(m.alpha.item(), float(bits_to_f32(ref["alpha_bits"])), rel_tol=2e-6)
        assert m.alpha.dim() == 0 and m.bias.abs().sum().item() == 0.0
    m = ob.QuantizedLinear(8, 4, bias=False)
    assert m.bias is None and list(m.state_dict().keys()) == ["weight", "alpha"]
    assert list(ob.QuantizedLinear(8, 4).state_dict().keys()) == ["weight", "alpha", "bias"]
    assert ob.BitLinear is ob.QuantizedLinear


def test_forward_interface_errors_and_fp32_bypass():
    torch.manual_seed(1)
    m = ob.QuantizedLinear(64, 64)
    x = torch.randn(2, 5, 64)
    y = m(x, 32)                                            # bitwidth 32 bypasses the quantiser (quant.py:121)
    assert torch.allclose(y, torch.nn.functional.linear(x, m.weight, m.bias))
    with pytest.raises(ValueError, match="bitwidth must be one of"):
        m(x, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, 2)                                             # the quantised path never runs on the CPU
    with pytest.raises(ValueError):
        ob.quantize_weight(m.weight, m.alpha, 3)
    assert ob.quantize_weight(m.weight, m.alpha, 32) is m.weight


@pytest.mark.skipif(not os.path.isdir("/root/reference/onebit_asr"), reason="reference tree not mounted")
def test_swaps_into_unmodified_reference_conformer():
    """The drop-in seam of SURVEY.md section 8(b): conformer.py does `from quant import QuantizedLinear`."""
    saved = {k: sys.modules.get(k) for k in ("quant", "conformer")}
    sys.path.insert(0, "/root/reference/onebit_asr")
    try:
        sys.modules.pop("conformer", None)
        ob.install_as_reference_quant()
        import conformer
        model = conformer.ConformerASR(80, 32, enc_layers=2, dec_layers=1)
        routed = [m for m in model.modules() if isinstance(m, ob.QuantizedLinear)]
        assert len(routed) == 2 * 9                          # 9 routed projections per block
        batch = {"feats": torch.randn(2, 64, 80), "feat_lens": torch.tensor([64, 48])}
        enc, mask, logits = model(batch, precision=32)       # fp32 bypass works on CPU
        assert enc.shape == (2, 15, 256) and logits.shape == (2, 15, 32)
    finally:
        sys.path.remove("/root/reference/onebit_asr")
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_gemm_operand_descriptors():
    """Shape/stride -> ob_gemm_f32 operand description (host logic of matmul.bmm_nt), on CPU tensors."""
    import onebit_b200  # noqa: F401
    from onebit_b200.matmul import _describe
    proj = torch.empty(3, 399, 256)
    heads = proj.view(3, 399, 4, 64).permute(0, 2, 1, 3)                       # [B, H, T, d] view of [B, T, H*d]
    assert _describe(heads.shape, heads.stride()) == (3, 4, 399, 64, 0, 256, 399 * 256, 64)
    t = heads.transpose(-1, -2)                                                # MN-major: the row axis is contiguous
    assert _describe(t.shape, t.stride()) == (3, 4, 64, 399, 1, 256, 399 * 256, 64)
    scores = torch.empty(3, 4, 399, 400)[..., :399]                            # padded pitch
    assert _describe(scores.shape, scores.stride())[4:6] == (0, 400)
    w = torch.empty(5004, 256)
    assert _describe(w.shape, w.stride()) == (1, 1, 5004, 256, 0, 256, 0, 0)
    assert _describe(w.t().shape, w.t().stride()) == (1, 1, 256, 5004, 1, 256, 0, 0)
    pos = torch.empty(1, 399, 256).view(1, 399, 4, 64).permute(0, 2, 1, 3)
    assert _describe(pos.shape, pos.stride())[6:] == (0, 64)                   # size-1 batch axis -> broadcast
    assert _describe((77, 45), (45, 1)) is None                                # pitch not a multiple of 4: packed copy
    assert _describe((8, 8, 8), (1, 8, 64)) is None or _describe((8, 8, 8), (1, 8, 64))[4] in (0, 1)
    with pytest.raises(ValueError):
        _describe((4,), (1,))


def test_route_audit_counts_and_strict_mode():
    """routes.py: torch routes are only counted (and, in strict mode, refused) for CUDA tensors; switched-off ops may pass."""
    import torch
    from onebit_b200 import routes
    from onebit_b200.training import reserve_allocator_headroom

    class OnDevice:                     # what the audit looks at
        is_cuda, dtype, shape = True, torch.float16, (2, 3)

    routes.reset()
    assert routes.taken("attention", False, torch.zeros(2)) is False            # CPU tensor (oracle-driven runs): not an event
    assert routes.taken("attention", True, OnDevice()) is True
    assert routes.taken("layer_norm", False, OnDevice()) is False
    assert routes.counts() == {"library": {"attention": 1}, "torch": {"layer_norm": 1}}
    routes.strict(True)
    try:
        with pytest.raises(RuntimeError, match="OB_STRICT_ROUTES"):
            routes.taken("layer_norm", False, OnDevice())
        assert routes.taken("conv_module", False, OnDevice(), switched_off=True) is False     # A/B switch: allowed
    finally:
        routes.strict(False)
        routes.reset()
    assert reserve_allocator_headroom("cpu", 6.0) == 0                           # nothing to reserve off the device
