"""GPU tests of the CTC loss kernels (csrc/ob_ctc.cu) through the C ABI: value and gradient against the reference fixture
(tests/golden/kat_ctc.npz, made by executing onebit_asr/losses.py:41-47), the numpy oracle, and torch's own CTC on the device
at the training shape.  Tolerances: loss 1e-5 relative; gradient 2e-5 absolute on the fixtures (|grad| <= 1), 1e-4 against the
fp64 oracle on longer inputs, and at the
training shape 1e-3 of the largest entry (both sides carry fp32 forward/backward variables of magnitude ~3e3, ulp 2.4e-4)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle import ctc_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctc():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import onebit_b200  # noqa: F401
    from onebit_b200 import ctc as mod
    return mod


def _run(ctc, logits, in_lens, targets, tgt_lens, blank, grad_out=1.0):
    x = torch.from_numpy(np.asarray(logits, dtype=np.float32)).cuda().requires_grad_(True)
    loss = ctc.ctc_loss(x, torch.from_numpy(np.asarray(in_lens)), torch.from_numpy(np.asarray(targets)).cuda(),
                        torch.from_numpy(np.asarray(tgt_lens)), blank)
    (grad,) = torch.autograd.grad(grad_out * loss, x)
    return loss.item(), grad.cpu().numpy()


@pytest.mark.parametrize("case", ["ragged", "repeats", "empty_and_infeasible", "longer"])
def test_ctc_reference_fixture(ctc, case):
    fx = np.load(os.path.join(GOLDEN, "kat_ctc.npz"))
    get = lambda k: fx[f"{case}.{k}"]  # noqa: E731
    loss, grad = _run(ctc, get("logits"), get("in_lens"), get("targets"), get("tgt_lens"), int(get("blank")), float(get("grad_out")))
    assert abs(loss - float(get("loss"))) <= 1e-5 * abs(float(get("loss")))
    assert np.abs(grad - get("grad")).max() <= 2e-5
    beyond = get("in_lens")[:, None] <= np.arange(get("logits").shape[1])[None, :]
    assert not np.any(grad[beyond])
    if case == "empty_and_infeasible":
        assert not np.any(grad[1])


@pytest.mark.parametrize("B,T,V,L,blank", [(1, 1, 5, 1, 0), (3, 17, 33, 6, 3), (2, 40, 1001, 19, 3), (5, 50, 64, 0, 3)])
def test_ctc_vs_oracle(ctc, B, T, V, L, blank):
    rng = np.random.default_rng(B * 1000 + T)
    logits = (2.0 * rng.standard_normal((B, T, V))).astype(np.float32)
    targets = rng.integers(0, V - 1, size=(B, L))
    targets = targets + (targets >= blank)
    if L >= 4:
        targets[:, 1] = targets[:, 0]                                  # repeated labels
    in_lens = rng.integers(max(1, T // 2), T + 1, size=B)
    in_lens[0] = T
    tgt_lens = rng.integers(0, L + 1, size=B) if L else np.zeros(B, dtype=np.int64)
    if L:
        tgt_lens[0] = L
    want_loss, want_grad, _ = ctc_oracle.ctc_loss_and_grad(logits, in_lens, targets, tgt_lens, blank, 2.0)
    loss, grad = _run(ctc, logits, in_lens, targets.reshape(B, L), tgt_lens, blank, 2.0)
    assert abs(loss - want_loss) <= 1e-5 * max(1.0, abs(want_loss))
    # fp32 forward/backward variables of magnitude ~250 (ulp 1.5e-5) against the fp64 oracle: occupancies carry ~1e-4 relative
    assert np.abs(grad - want_grad).max() <= 1e-4


def test_ctc_training_shape_vs_torch(ctc):
    """64 x 399 frames x 5004 classes, 64 labels per utterance (the bench workload): against F.ctc_loss on the device."""
    g = torch.Generator().manual_seed(7)
    B, T, V, L, blank = 64, 399, 5004, 64, 3
    x = torch.randn(B, T, V, generator=g).cuda().requires_grad_(True)
    targets = torch.randint(4, V, (B, L), generator=g).cuda()
    in_lens = torch.randint(300, T + 1, (B,), generator=g)
    in_lens[0] = T
    tgt_lens = torch.randint(1, L + 1, (B,), generator=g)
    loss = ctc.ctc_loss(x, in_lens.cuda(), targets, tgt_lens.cuda(), blank)
    (grad,) = torch.autograd.grad(loss, x)
    x2 = x.detach().clone().requires_grad_(True)
    ref = F.ctc_loss(F.log_softmax(x2, dim=-1).transpose(0, 1), targets, in_lens, tgt_lens, blank=blank, reduction="mean",
                     zero_infinity=True)
    (ref_grad,) = torch.autograd.grad(ref, x2)
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (grad - ref_grad).abs().max().item() <= 1e-3 * ref_grad.abs().max().item()
    # size-independent property: every valid frame's gradient sums to zero over the classes (softmax minus a distribution)
    rows = grad.sum(dim=-1).abs().max().item()
    assert rows <= 1e-3 * grad.abs().max().item() * 10
    # deterministic: same bits on a second evaluation
    loss_b = ctc.ctc_loss(x, in_lens.cuda(), targets, tgt_lens.cuda(), blank)
    (grad_b,) = torch.autograd.grad(loss_b, x)
    assert torch.equal(loss, loss_b) and torch.equal(grad, grad_b)


def test_ctc_argument_errors(ctc):
    x = torch.zeros(2, 4, 8, device="cuda")
    with pytest.raises(ValueError):
        ctc.ctc_loss(x, torch.tensor([4, 4]), torch.zeros(2, 3, dtype=torch.long, device="cuda"), torch.tensor([1, 1]), 8)   # blank >= V
    with pytest.raises(ValueError):
        ctc.ctc_loss(x, torch.tensor([4]), torch.zeros(2, 3, dtype=torch.long, device="cuda"), torch.tensor([1, 1]), 0)      # lens != B
    with pytest.raises(RuntimeError):
        ctc.ctc_loss(x.cpu(), torch.tensor([4, 4]), torch.zeros(2, 3, dtype=torch.long), torch.tensor([1, 1]), 0)             # no fallback
