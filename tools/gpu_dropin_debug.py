"""Where does the unmodified reference conformer on the B200 layer depart from the Oracle-B layer?  (debug aid)"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import onebit_b200 as ob
from oracle import ref_loader
from oracle.torch_oracle import OracleQuantizedLinear

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

class Layer(OracleQuantizedLinear):
    act_bits_default = 8
mod = types.ModuleType("quant"); mod.QuantizedLinear = Layer
oracle, product = ref_loader.load(quant_module=mod), ref_loader.load(quant_module=ob.quant)
kw = dict(enc_layers=3, dec_layers=1, enc_dropout=0.0, dec_dropout=0.0)
torch.manual_seed(21); m_o = oracle.conformer.ConformerASR(80, 64, **kw).train()
torch.manual_seed(21); m_p = product.conformer.ConformerASR(80, 64, **kw).train().cuda()
g = torch.Generator().manual_seed(77)
batch = {"feats": torch.randn(3, 131, 80, generator=g), "feat_lens": torch.tensor([131, 100, 64])}
bg = {k: v.cuda() for k, v in batch.items()}
acts = {}
def hook(tag):
    def f(m, i, o):
        acts.setdefault(tag, {})[m._dbg] = (o[0] if isinstance(o, tuple) else o).detach().float().cpu()
    return f
for tag, m in (("cpu", m_o), ("gpu", m_p)):
    for n, sub in m.named_modules():
        if n.count(".") <= 3 and n:
            sub._dbg = n
            sub.register_forward_hook(hook(tag))
with torch.no_grad():
    m_o(batch, precision=2)
    m_p(bg, precision=2)
    m_og = m_o.cuda()
    for n, sub in m_og.named_modules():
        pass
    acts["ogpu"] = {}
    def hook2(m, i, o):
        acts["ogpu"][m._dbg] = (o[0] if isinstance(o, tuple) else o).detach().float().cpu()
    hs = [sub.register_forward_hook(hook2) for n, sub in m_og.named_modules() if hasattr(sub, "_dbg")]
    m_og(bg, precision=2)
for n in acts["cpu"]:
    a, b, c = acts["cpu"][n], acts["gpu"].get(n), acts["ogpu"].get(n)
    if b is None or a.shape != b.shape:
        continue
    d = lambda u, v: float((u - v).abs().max() / (v.abs().max() + 1e-30))
    print(f"{n:40s} product-vs-cpuoracle {d(b, a):.2e}   gpuoracle-vs-cpuoracle {d(c, a):.2e}   product-vs-gpuoracle {d(b, c):.2e}")
