"""CPU checks of the Conformer caller and the co-training step against the reference-generated fixture
(tests/golden/make_golden_conformer.py), plus the data-parallel gradient exchange on gloo (world_size 2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT

import onebit_b200 as ob
from onebit_b200.training import StepConfig, WarmupCosine, cotraining_loss, sample_sp_mask
from oracle.torch_oracle import OracleQuantizedLinear

CFG = dict(input_dim=80, vocab_size=64, enc_layers=3, dec_layers=1, enc_dropout=0.0, dec_dropout=0.0)


@pytest.fixture(scope="module")
def fx():
    return dict(np.load(os.path.join(GOLDEN, "conformer_step.npz")))


def load_batch(fx, device="cpu"):
    b = {k: torch.from_numpy(fx[k]).to(device) for k in ("feats", "feat_lens", "tokens", "token_lens")}
    b["feat_lens_cpu"] = torch.from_numpy(fx["feat_lens"])
    b["token_lens_cpu"] = torch.from_numpy(fx["token_lens"])
    return b


def build(act_bits, seed):
    torch.manual_seed(seed)
    OracleQuantizedLinear.act_bits_default = act_bits
    try:
        return ob.ConformerASR(**CFG, linear_cls=OracleQuantizedLinear).train()
    finally:
        OracleQuantizedLinear.act_bits_default = 8


def run_steps(model, batch, sp_masks, lr, cfg=StepConfig()):
    opt = torch.optim.AdamW(model.parameters(), lr=lr, betas=(0.9, 0.98), weight_decay=1e-2)
    out = []
    for spm in sp_masks:
        opt.zero_grad()
        loss, _ = cotraining_loss(model, batch, cfg, list(spm))
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
        opt.step()
        out.append(loss.item())
    return out


def test_training_run_reproduces_reference_losses(fx):
    """Oracle-A through our module tree == the reference's run_epoch arithmetic, 4 optimiser steps, batch 3."""
    torch.set_num_threads(1)
    model = build(32, int(fx["seed"]))
    losses = run_steps(model, load_batch(fx), fx["sp_masks"], float(fx["lr"]))
    np.testing.assert_allclose(losses, fx["losses_A"], rtol=2e-4)


def test_host_lengths_and_shared_frontend_do_not_change_the_loss(fx):
    torch.set_num_threads(1)
    batch = load_batch(fx)
    plain = {k: v for k, v in batch.items() if not k.endswith("_cpu")}
    m = build(8, 3)
    l0, _ = cotraining_loss(m, plain, StepConfig(), [1, 0, 1])
    l1, _ = cotraining_loss(m, batch, StepConfig(), [1, 0, 1])
    l2, _ = cotraining_loss(m, batch, StepConfig(share_frontend=True), [1, 0, 1])
    assert abs(l0.item() - l1.item()) < 1e-6 and abs(l0.item() - l2.item()) < 1e-5
    g_ref = torch.autograd.grad(l1, m.encoder.subsample.out.weight, retain_graph=True)[0]
    g_shared = torch.autograd.grad(l2, m.encoder.subsample.out.weight)[0]
    assert torch.allclose(g_ref, g_shared, rtol=1e-4, atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/onebit_asr"), reason="reference tree not mounted")
@pytest.mark.parametrize("sp_mask", [[1, 0, 1], [0, 0, 0], [1, 1, 1]])
def test_stacked_passes_equal_sequential_passes(fx, sp_mask):
    """The three co-training passes evaluated side by side on one stacked batch (StepConfig.stack_passes) give the loss and
    the gradients of three separate passes: per-bitwidth row groups in the routed layers, per-pass BatchNorm statistics."""
    torch.set_num_threads(1)
    batch = load_batch(fx)
    m = build(8, 3)
    results = []
    for cfg in (StepConfig(share_frontend=True), StepConfig(share_frontend=True, stack_passes=True)):
        m.zero_grad(set_to_none=True)
        loss, parts = cotraining_loss(m, batch, cfg, sp_mask)
        loss.backward()
        results.append((loss.item(), {k: v.item() for k, v in parts.items()},
                        {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}))
    (l_seq, p_seq, g_seq), (l_stk, p_stk, g_stk) = results
    assert abs(l_seq - l_stk) < 1e-5 * abs(l_seq)
    for k in p_seq:
        assert abs(p_seq[k] - p_stk[k]) < 1e-5 * max(1.0, abs(p_seq[k])), k
    assert g_seq.keys() == g_stk.keys()
    for n in g_seq:
        scale = g_seq[n].abs().max().item() + 1e-12
        assert (g_seq[n] - g_stk[n]).abs().max().item() < 2e-4 * scale + 1e-7, n


def test_stack_plan():
    from onebit_b200.conformer import ConformerEncoder
    plan = ConformerEncoder.stack_plan
    assert plan([(2, None), (1, None), (2, [1, 0])]) == ([0, 2, 1], 1, [1, 0])
    assert plan([(1, None), (2, None)]) == ([1, 0], 1, None)
    assert plan([(2, None), (32, None)]) is None                      # full precision passes are not stacked
    assert plan([(2, [0]), (2, [1])]) is None                        # two stochastic passes: 2-bit rows would not be a prefix
    assert plan([(2, None)]) is None


def test_written_out_decoder_equals_torch_decoder(monkeypatch):
    """TransformerDecoder._layers_on_library (every projection through matmul.linear) is nn.TransformerDecoder's computation:
    forced onto CPU with F.linear standing in for the GEMM, outputs and gradients must agree with the torch module."""
    from onebit_b200 import matmul
    from onebit_b200.conformer import TransformerDecoder
    torch.manual_seed(0)
    dec = TransformerDecoder(64, 256, 2, 4, 1024, 0.0, 0).train()
    tgt = torch.randint(1, 64, (3, 9))
    tgt[1, 6:] = 0
    tgt[2, 4:] = 0
    mem = torch.randn(3, 21, 256)
    mem_mask = torch.ones(3, 21, dtype=torch.bool)
    mem_mask[1, 15:] = False
    mem_mask[2, 5:] = False

    def run():
        dec.zero_grad()
        m = mem.clone().requires_grad_(True)
        y = dec(tgt, m, mem_mask, tgt == 0)
        y.square().sum().backward()
        return y.detach(), m.grad, {n: p.grad.clone() for n, p in dec.named_parameters()}

    y_ref, gm_ref, gp_ref = run()
    monkeypatch.setattr(matmul, "linear_usable", lambda x, w: True)
    monkeypatch.setattr(matmul, "linear", lambda x, w, b=None: torch.nn.functional.linear(x, w, b))
    y_lib, gm_lib, gp_lib = run()
    assert (y_ref - y_lib).abs().max().item() < 1e-6
    assert (gm_ref - gm_lib).abs().max().item() < 1e-5 * gm_ref.abs().max().item()
    for n in gp_ref:
        assert (gp_ref[n] - gp_lib[n]).abs().max().item() < 1e-5 * gp_ref[n].abs().max().item() + 1e-9, n


def test_module_tree_matches_reference_state_dict():
    saved = {k: sys.modules.get(k) for k in ("quant", "conformer")}
    sys.path.insert(0, "/root/reference/onebit_asr")
    try:
        for k in ("quant", "conformer"):
            sys.modules.pop(k, None)
        import conformer as refc
        torch.manual_seed(5)
        ref = refc.ConformerASR(80, 40, enc_layers=2, dec_layers=1)
        torch.manual_seed(5)
        ours = ob.ConformerASR(80, 40, enc_layers=2, dec_layers=1)
        sd_r, sd_o = ref.state_dict(), ours.state_dict()
        assert list(sd_r.keys()) == list(sd_o.keys())
        assert all(torch.equal(sd_r[k], sd_o[k]) for k in sd_r)
        ours.load_state_dict(sd_r)                                 # reference checkpoints load (eval.py:282)
        batch = {"feats": torch.randn(2, 90, 80), "feat_lens": torch.tensor([90, 50])}
        ref.eval(), ours.eval()
        with torch.no_grad():
            a, b = ref(batch, precision=32), ours(batch, precision=32)
        assert torch.allclose(a[0], b[0], atol=1e-6) and torch.equal(a[1], b[1]) and torch.allclose(a[2], b[2], atol=1e-5)
    finally:
        sys.path.remove("/root/reference/onebit_asr")
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_schedule_helpers():
    torch.manual_seed(0)
    m = sample_sp_mask(12)
    assert len(m) == 12 and set(m) <= {0, 1}
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=5e-4)
    s = WarmupCosine(opt, warmup_steps=4, total_steps=12)
    lrs = []
    for _ in range(12):
        s.step()
        lrs.append(opt.param_groups[0]["lr"])
    assert abs(lrs[0] - 5e-4 / 4) < 1e-12 and abs(max(lrs) - 5e-4) < 1e-12 and abs(lrs[-1] - 5e-5) < 1e-12


# ------------------------------------------------------------------------------------------ data parallel, gloo x2
def _dp_worker(rank, world, port, fxpath, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    fx = dict(np.load(fxpath))
    batch = load_batch(fx)
    # rank r takes utterance r (ragged lengths): the two ranks see different data
    shard = {k: (v[rank:rank + 1] if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in batch.items()}
    model = build(8, 3)
    from onebit_b200.dp import GradAllReducer
    sync = GradAllReducer(model.parameters(), bucket_bytes=1 << 20)
    loss, _ = cotraining_loss(model, shard, StepConfig(), [1, 0, 1])
    loss.backward()
    sync.finish()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    if rank == 0:
        q.put(flat.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_allreduce_gloo(fx):
    """mean of per-rank gradients (overlapped bucketed all-reduce) == gradients of the averaged per-rank losses."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, os.path.join(GOLDEN, "conformer_step.npz"), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.set_num_threads(1)
    batch = load_batch(fx)
    model = build(8, 3)
    total = 0.0
    for r in range(2):
        shard = {k: (v[r:r + 1] if torch.is_tensor(v) and v.dim() > 0 else v) for k, v in batch.items()}
        loss, _ = cotraining_loss(model, shard, StepConfig(), [1, 0, 1])
        total = total + 0.5 * loss
    total.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).numpy()
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-6)
