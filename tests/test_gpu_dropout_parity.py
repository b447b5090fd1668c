"""End-to-end parity UNDER DROPOUT (p = 0.1, what bench.py times): the encoder on the B200 kernels - fused modules, dropout masks
generated inside the kernels from Philox streams - against the Oracle-B module tree on the CPU fed with the SAME masks.

The product draws one (seed, offset, threshold) stream per dropout site (quant.draw_dropout_stream); the test records them in
call order and the oracle run replaces every ``nn.Dropout`` of the encoder blocks by a module that rebuilds the mask of the next
stream with the oracle's Philox restatement (oracle/onebit_oracle.py: ``dropout_keep_flat`` for the FFN mid-section,
``dropout_keep_groups8`` for the module tails, ``dropout_keep_relattn`` for the attention weights).  The positional-encoding
dropout (torch's own RNG on either device) is switched off on both sides.

tolerances: first feed-forward module rel <= 1e-2 of max (identical masks; int8 code flips at the two quantisers - the front-end
            runs on different devices - separate the two; measured 4.5e-3; a wrong mask gives O(1)),
            encoder output rel <= 3e-2 (flip amplification, see tests/test_reference_dropin.py), gradient norms rel <= 3e-2.
"""
import numpy as np
import pytest
import torch

from oracle import onebit_oracle as orc

pytestmark = pytest.mark.gpu

P = 0.1


class _InjectedDropout(torch.nn.Module):
    """Applies the mask of the next recorded stream; the site kind follows from the tensor it is applied to."""

    def __init__(self, streams, d_ff):
        super().__init__()
        self.streams, self.d_ff, self.p = streams, d_ff, P

    def forward(self, x):
        if not self.training:
            return x
        inv_keep, (seed, offset, thr) = self.streams.pop(0)
        if x.dim() == 4:                                              # attention weights [B, H, T, T]
            keep = orc.dropout_keep_relattn(x.shape[0], x.shape[1], x.shape[2], seed, offset, thr)
        elif x.shape[-1] == self.d_ff:                                # FFN mid-section [B, T, d_ff]
            keep = orc.dropout_keep_flat(x.numel(), seed, offset, thr)
        else:                                                         # module tail [B, T, C]
            keep = orc.dropout_keep_groups8(x.numel(), seed, offset, thr)
        keep = torch.from_numpy(np.ascontiguousarray(keep).astype(np.float32)).reshape(x.shape)
        return x * keep * np.float32(inv_keep)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


@pytest.mark.parametrize("precision", [2, 1])
def test_encoder_under_dropout_matches_oracle_with_the_same_masks(precision):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import onebit_b200 as ob
    from onebit_b200 import quant as obq
    from oracle.torch_oracle import OracleQuantizedLinear
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(input_dim=80, vocab_size=64, enc_layers=2, dec_layers=1, enc_dropout=P, dec_dropout=0.0)
    torch.manual_seed(4)
    m_gpu = ob.ConformerASR(**cfg).train()
    torch.manual_seed(4)
    OracleQuantizedLinear.act_bits_default = 8
    m_cpu = ob.ConformerASR(**cfg, linear_cls=OracleQuantizedLinear).train()
    assert all(torch.equal(v, m_gpu.state_dict()[k]) for k, v in m_cpu.state_dict().items())
    m_gpu = m_gpu.cuda()
    for m in (m_gpu, m_cpu):
        m.encoder.pos_enc.dropout.p = 0.0
    g = torch.Generator().manual_seed(9)
    feats, lens = torch.randn(3, 259, 80, generator=g), torch.tensor([259, 200, 131])      # T_sub = 64: flat masks need n % 256 == 0

    streams, orig = [], obq.draw_dropout_stream

    def recording(device, p):
        out = orig(device, p)
        streams.append(out)
        return out
    first = {}
    hooks = [m.encoder.blocks[0].ff1.register_forward_hook(lambda mod, i, o, k=k: first.__setitem__(k, o.detach().float().cpu()))
             for k, m in (("gpu", m_gpu), ("cpu", m_cpu))]
    import onebit_b200.attention as att
    import onebit_b200.fused as fused
    import onebit_b200.residual as residual
    patched = [(mod, "draw_dropout_stream") for mod in (obq, att, fused, residual)]
    for mod, name in patched:
        setattr(mod, name, recording)
    try:
        torch.cuda.manual_seed(123)
        enc_g, valid_g = m_gpu.encoder(feats.cuda(), lens.cuda(), precision)
        w = torch.randn(enc_g.shape, generator=g)
        (enc_g * w.cuda()).sum().backward()
    finally:
        for mod, name in patched:
            setattr(mod, name, orig)
    assert len(streams) == 2 * 7                      # per block: ff1 (mid, tail), mhsa (weights, tail), conv (tail), ff2 (mid, tail)
    replay = list(streams)
    for blk in m_cpu.encoder.blocks:
        for mod in (blk.ff1, blk.mhsa, blk.conv, blk.ff2):
            mod.dropout = _InjectedDropout(replay, 1024)
    enc_c, valid_c = m_cpu.encoder(feats, lens, precision)
    (enc_c * w).sum().backward()
    assert not replay                                  # every recorded stream was consumed, in the same order
    for h in hooks:
        h.remove()
    assert torch.equal(valid_c, valid_g.cpu())
    assert _rel(first["gpu"], first["cpu"]) < 1e-2
    assert _rel(enc_g.detach().cpu(), enc_c.detach()) < 3e-2
    # dropout really happened, and the same one on both sides: a run without it differs by far more than the tolerance
    m_cpu.eval()
    with torch.no_grad():
        enc_nodrop, _ = m_cpu.encoder(feats, lens, precision)
    assert _rel(enc_nodrop, enc_c.detach()) > 0.2
    grads_c = {n: p.grad for n, p in m_cpu.encoder.named_parameters() if p.grad is not None}
    total = float(torch.sqrt(sum(v.double().pow(2).sum() for v in grads_c.values())))
    for n, p in m_gpu.encoder.named_parameters():
        if n not in grads_c:
            assert p.grad is None, n
            continue
        atol = 2e-3 * total if n.endswith(".alpha") else 1e-4 * total
        got, want = p.grad.double().norm().item(), grads_c[n].double().norm().item()
        assert abs(got - want) <= 3e-2 * want + atol, (n, got, want)
