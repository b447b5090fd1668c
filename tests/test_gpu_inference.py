"""GPU tests of the inference side: device greedy CTC decode vs the reference fixture / oracle, packed layers."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import onebit_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import onebit_b200
    return onebit_b200


def test_greedy_decode_reference_fixture(ob):
    fx = dict(np.load(os.path.join(GOLDEN, "kat_decode.npz")))
    toks, n = ob.ctc_greedy_decode(torch.from_numpy(fx["logits"]).cuda(), torch.from_numpy(fx["lens"]).cuda(), blank_id=3)
    assert np.array_equal(n.cpu().numpy(), fx["out_lens"])
    assert np.array_equal(toks.cpu().numpy(), fx["tokens"])          # index-exact, -1 padding included


@pytest.mark.parametrize("B,T,V,dtype", [(1, 1, 8, torch.float32), (5, 249, 5004, torch.float32),
                                          (64, 399, 5004, torch.float32), (7, 333, 1001, torch.bfloat16)])
def test_greedy_decode_vs_oracle(ob, B, T, V, dtype):
    g = torch.Generator().manual_seed(B * T + V)
    logits = torch.randn(B, T, V, generator=g)
    steer = torch.randint(0, 12, (B, T), generator=g)                 # repeats and blanks (id 3)
    logits.scatter_(2, steer.unsqueeze(-1), 7.0)
    logits = logits.to(dtype)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    from onebit_b200.inference import ctc_greedy_decode_lists
    got = ctc_greedy_decode_lists(logits.cuda(), lens.cuda(), blank_id=3)
    ref = orc.ctc_greedy_decode_batch(logits.float().numpy(), lens.numpy(), blank_id=3)
    assert got == ref


def test_packed_layer_equals_training_layer(ob):
    torch.manual_seed(3)
    layer = ob.QuantizedLinear(256, 1024).cuda()
    with torch.no_grad():
        layer.bias.normal_(0, 0.1)
    x = torch.randn(4, 99, 256, device="cuda")
    for bw in (1, 2):
        packed = ob.PackedQuantizedLinear.from_layer(layer, bw)
        with torch.no_grad():
            assert torch.equal(packed(x, bw), layer(x, bw))           # same kernels, same codes: bit-identical
        with pytest.raises(ValueError):
            packed(x, 3 - bw)
    assert sum(t.numel() * t.element_size() for t in packed.state_dict().values()) < 256 * 1024 * 4 / 10


def test_packed_model_roundtrip_and_transcribe(ob):
    from onebit_b200.inference import load_packed_state_dict, packed_state_dict, transcribe_greedy
    cfg = dict(input_dim=80, vocab_size=64, enc_layers=2, dec_layers=1, enc_dropout=0.0, dec_dropout=0.0)
    torch.manual_seed(4)
    torch.backends.cudnn.deterministic = True         # torch's front-end convolutions: same algorithm on every call
    model = ob.ConformerASR(**cfg).cuda().eval()
    batch = {"feats": torch.randn(3, 200, 80, device="cuda"), "feat_lens": torch.tensor([200, 150, 90], device="cuda")}
    toks_ref, n_ref = transcribe_greedy(model, batch, precision=2)
    state = packed_state_dict(model, 2)
    assert not any(k.endswith("lin1.weight") for k in state) and any(k.endswith("lin1.packed") for k in state)
    torch.manual_seed(99)
    fresh = ob.ConformerASR(**cfg).cuda()
    load_packed_state_dict(fresh, state)
    toks, n = transcribe_greedy(fresh, batch, precision=2)
    assert torch.equal(n, n_ref) and torch.equal(toks, toks_ref)
    assert n.tolist() == [min(int(x), 49) for x in n.tolist()] and toks.shape == (3, 49)


def test_graphed_transcriber_matches_eager(ob):
    from onebit_b200.inference import GraphedTranscriber, pack_model_for_inference, transcribe_greedy
    cfg = dict(input_dim=80, vocab_size=64, enc_layers=2, dec_layers=1, enc_dropout=0.0, dec_dropout=0.0)
    torch.manual_seed(8)
    model = pack_model_for_inference(ob.ConformerASR(**cfg).cuda(), 2)
    runner = GraphedTranscriber(model, batch_size=3, frames=200, precision=2)
    for seed in (1, 2):
        g = torch.Generator().manual_seed(seed)
        feats = torch.randn(3, 200, 80, generator=g).cuda()
        lens = torch.tensor([200, 120 + seed, 64]).cuda()
        toks, n = runner(feats, lens)
        toks_e, n_e = transcribe_greedy(model, {"feats": feats, "feat_lens": lens}, precision=2)
        assert torch.equal(n, n_e) and torch.equal(toks, toks_e)
