"""Pin the CPU oracle (oracle/) against the golden fixtures produced by executing the
reference (tests/golden/make_golden.py).  CPU only."""
import hashlib
import math

import numpy as np
import pytest
import torch

from conftest import bits_to_f32
from oracle import onebit_oracle as ob
from oracle.torch_oracle import OracleQuantizedLinear


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("bw", [1, 2])
@pytest.mark.parametrize("case", ["", "alpha2_"])
def test_edges_bit_exact(kat_edges, bw, case):
    W = np.array(kat_edges["W2" if case else "W"], dtype=np.float32)
    a = np.float32(kat_edges["alpha2" if case else "alpha"])
    g = np.array(kat_edges["g"], dtype=np.float32)
    ref = kat_edges[f"{case}bw{bw}"]
    assert np.array_equal(ob.quantize_weight(W, a, bw), np.array(ref["W_hat"], dtype=np.float32))
    gw, ga = ob.ste_backward(g, W, a, bw)
    assert np.array_equal(gw, np.array(ref["grad_W"], dtype=np.float32))
    assert math.isclose(float(ga), ref["grad_alpha"], rel_tol=1e-6, abs_tol=1e-6)


def test_seeded_codes_hashes(kat_seeded):
    for key, ref in kat_seeded.items():
        torch.manual_seed(0)
        w, a, b = ob.init_layer_params(ref["in"], ref["out"])
        W = w.numpy()
        assert sha16(W) == ref["sha_W"], key                      # same RNG stream as the reference ctor
        assert math.isclose(float(a), float(bits_to_f32(ref["alpha_bits"])), rel_tol=2e-6)
        a_eff = bits_to_f32(ref["alpha_eff_bits"])
        assert ob.alpha_eff(bits_to_f32(ref["alpha_bits"])) == a_eff
        q2, q1 = ob.quant_codes(W, a_eff, 2), ob.quant_codes(W, a_eff, 1)
        assert sha16(q2) == ref["sha_q2"] and sha16(q1) == ref["sha_q1"], key
        assert int(q2.astype(np.int64).sum()) == ref["sum_q2"] and int((q2 != 0).sum()) == ref["nnz_q2"]
        assert int(ob.ste_mask(W, a_eff).sum()) == ref["ste_in_window"]


def test_act_quant_bit_exact(kat_layer, kat_layer_stats):
    q, s = ob.act_quant(kat_layer["x"])
    assert np.array_equal(q, kat_layer["act_q"]) and np.array_equal(s, kat_layer["act_s"])
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 249, 256, generator=g).numpy()
    q, s = ob.act_quant(x)
    assert sha16(q) == kat_layer_stats["act"]["sha_q"] and sha16(s) == kat_layer_stats["act"]["sha_s"]


@pytest.mark.parametrize("bw", [1, 2, 32])
@pytest.mark.parametrize("oracle", ["A", "B"])
def test_layer_forward_backward(kat_layer, bw, oracle):
    if bw == 32 and oracle == "B":
        pytest.skip("bitwidth 32 bypasses both quantisers")
    k = kat_layer
    act_bits = 8 if oracle == "B" else 32
    tag = f"bw{bw}_{oracle}"
    y = ob.linear_forward(k["x"], k["W"], k["alpha"], k["bias"], bw, act_bits)
    np.testing.assert_allclose(y, k[f"{tag}_y"], rtol=1e-5, atol=1e-5)
    g = ob.linear_backward(k["gy"], k["x"], k["W"], k["alpha"], k["bias"], bw, act_bits)
    np.testing.assert_allclose(g["x"], k[f"{tag}_gx"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(g["weight"], k[f"{tag}_gW"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(g["bias"], k[f"{tag}_gb"], rtol=1e-5, atol=1e-5)
    if bw != 32:
        assert (g["weight"] != 0).sum() == (k[f"{tag}_gW"] != 0).sum()      # STE mask is exact
        assert math.isclose(float(g["alpha"]), float(k[f"{tag}_galpha"]), rel_tol=2e-5, abs_tol=1e-3)
        assert np.array_equal(ob.quant_codes(k["W"], ob.alpha_eff(k["alpha"]), bw), k[f"codes_bw{bw}"])


@pytest.mark.parametrize("bw", [1, 2, 32])
@pytest.mark.parametrize("oracle", ["A", "B"])
def test_torch_oracle_layer(kat_layer, bw, oracle):
    if bw == 32 and oracle == "B":
        pytest.skip("bitwidth 32 bypasses both quantisers")
    k = kat_layer
    tag = f"bw{bw}_{oracle}"
    m = OracleQuantizedLinear(128, 192, act_bits=8 if oracle == "B" else 32)
    with torch.no_grad():
        m.weight.copy_(torch.from_numpy(k["W"]))
        m.alpha.copy_(torch.tensor(float(k["alpha"])))
        m.bias.copy_(torch.from_numpy(k["bias"]))
    x = torch.from_numpy(k["x"]).requires_grad_(True)
    y = m(x, bw)
    y.backward(torch.from_numpy(k["gy"]))
    np.testing.assert_allclose(y.detach().numpy(), k[f"{tag}_y"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), k[f"{tag}_gx"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(m.weight.grad.numpy(), k[f"{tag}_gW"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(m.bias.grad.numpy(), k[f"{tag}_gb"], rtol=1e-6, atol=1e-6)
    if bw != 32:
        assert math.isclose(m.alpha.grad.item(), float(k[f"{tag}_galpha"]), rel_tol=1e-5, abs_tol=1e-4)
    with pytest.raises(ValueError):
        m(x, 4)


def test_layer_stats(kat_layer_stats):
    torch.manual_seed(0)
    w, a, b = ob.init_layer_params(256, 1024)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 249, 256, generator=g).numpy()
    gy = torch.randn(4, 249, 1024, generator=g).numpy()
    for bw in (1, 2):
        for oracle, act_bits in (("A", 32), ("B", 8)):
            ref = kat_layer_stats[f"bw{bw}_{oracle}"]
            y = ob.linear_forward(x, w.numpy(), a.numpy(), b.numpy(), bw, act_bits)
            gr = ob.linear_backward(gy, x, w.numpy(), a.numpy(), b.numpy(), bw, act_bits)
            assert math.isclose(float(y.astype(np.float64).sum()), ref["sum_y"], rel_tol=1e-4)
            assert math.isclose(float(np.abs(gr["x"]).mean()), ref["mean_abs_gx"], rel_tol=1e-5)
            assert math.isclose(float(np.abs(gr["weight"]).mean()), ref["mean_abs_gW"], rel_tol=1e-5)
            assert math.isclose(float((gr["weight"] != 0).mean()), ref["nnz_frac_gW"], rel_tol=1e-9)
            assert math.isclose(float(gr["alpha"]), ref["galpha"], rel_tol=1e-4)


@pytest.mark.parametrize("order", ["i8", "bf16"])
def test_pack_roundtrip(order):
    rng = np.random.default_rng(3)
    Q = rng.integers(-1, 2, size=(37, 128)).astype(np.int8)
    P = ob.pack_codes(Q, order)
    assert P.shape == (37, 32) and P.dtype == np.uint8
    assert np.array_equal(ob.unpack_codes(P, order), Q)
    # zero bytes decode to zero codes (TMA out-of-bounds fill is therefore harmless)
    assert not ob.unpack_codes(np.zeros((2, 8), np.uint8), order).any()


def test_greedy_decode_matches_reference_fixture():
    import os
    from conftest import GOLDEN
    fx = dict(np.load(os.path.join(GOLDEN, "kat_decode.npz")))
    hyps = ob.ctc_greedy_decode_batch(fx["logits"], fx["lens"], blank_id=3)
    for b, h in enumerate(hyps):
        assert len(h) == fx["out_lens"][b] and h == fx["tokens"][b, : len(h)].tolist()


# ------------------------------------------------------------------ dropout stream of the fused kernels
def test_philox_known_answers():
    """Philox4x32-10 restatement against the Random123 known-answer vectors (counter, key -> output)."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = ob.philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(w[0]) for w in got) == want


def test_dropout_masks_are_bernoulli_and_keyed():
    thr = int(round(0.1 * 2 ** 16))
    a = ob.dropout_keep_flat(1 << 18, seed=7, offset=0, threshold=thr)
    b = ob.dropout_keep_flat(1 << 18, seed=7, offset=4, threshold=thr)
    assert abs(a.mean() - 0.9) < 5e-3 and abs(b.mean() - 0.9) < 5e-3
    assert 0.75 < (a == b).mean() < 0.89                      # independent streams agree on 0.81 + 0.01 of the elements
    k = ob.dropout_keep_relattn(2, 3, 70, seed=7, offset=0, threshold=thr)
    assert k.shape == (2, 3, 70, 70) and abs(k.mean() - 0.9) < 1e-2
    assert ob.dropout_keep_flat(256, 1, 0, 0).all()            # threshold 0 keeps everything


@pytest.mark.parametrize("case", ["ragged", "repeats", "empty_and_infeasible", "longer"])
def test_ctc_oracle_matches_reference_fixture(case):
    """oracle/ctc_oracle.py against the reference's own ctc_loss_from_logits (tests/golden/make_golden_ctc.py)."""
    import os

    from conftest import GOLDEN
    from oracle import ctc_oracle

    fx = np.load(os.path.join(GOLDEN, "kat_ctc.npz"))
    get = lambda k: fx[f"{case}.{k}"]  # noqa: E731
    loss, grad, nll = ctc_oracle.ctc_loss_and_grad(get("logits"), get("in_lens"), get("targets"), get("tgt_lens"), int(get("blank")),
                                                   float(get("grad_out")))
    assert abs(loss - float(get("loss"))) <= 1e-6 * abs(float(get("loss")))          # fp32 reference vs fp64 restatement
    assert np.abs(grad - get("grad")).max() <= 2e-5
    if case == "empty_and_infeasible":                                             # zero_infinity: no loss, no gradient
        assert np.isinf(nll[1]) and not np.any(get("grad")[1]) and not np.any(grad[1])
    beyond = get("in_lens")[:, None] <= np.arange(get("logits").shape[1])[None, :]
    assert not np.any(grad[beyond])                                                # frames past the input length
