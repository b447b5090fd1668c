"""Data parallelism on hardware: two B200s, NCCL.  Skipped on a single-GPU box; run with
``gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu``.

DP-2 gradients (bucketed all-reduce overlapped with backward, ``.grad`` left as views of the buckets) must equal the
gradients one GPU accumulates over the same two micro-batches - identical math, because BatchNorm uses per-replica batch
statistics in both (conformer.py:148).  Tolerance rtol 1e-4 (fp32 summation order only: the per-micro-batch kernels are the
same on both sides and the 1/2 weighting is exact in bf16).  Also checked: rank 1 starts from DIFFERENT weights and is brought
in line by the constructor's broadcast; an optimiser step on the bucket-view gradients leaves both ranks with equal weights.
"""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

CFG = dict(input_dim=80, vocab_size=64, enc_layers=3, dec_layers=1, enc_dropout=0.0, dec_dropout=0.0)
SP_MASK = [1, 0, 1]


def _shard(fx, r, device):
    b = {k: torch.from_numpy(fx[k])[r:r + 1].to(device) for k in ("feats", "feat_lens", "tokens", "token_lens")}
    b["feat_lens_cpu"] = torch.from_numpy(fx["feat_lens"])[r:r + 1]
    b["token_lens_cpu"] = torch.from_numpy(fx["token_lens"])[r:r + 1]
    return b


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import onebit_b200 as ob
    from onebit_b200.dp import GradAllReducer
    from onebit_b200.training import StepConfig, cotraining_loss
    fx = dict(np.load(os.path.join(GOLDEN, "conformer_step.npz")))
    torch.manual_seed(3 + 100 * rank)                       # rank 1 is initialised differently on purpose
    model = ob.ConformerASR(**CFG).train().cuda()
    sync = GradAllReducer(model.parameters(), bucket_bytes=1 << 20, buffers=model.buffers())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    opt.zero_grad(set_to_none=True)
    loss, _ = cotraining_loss(model, _shard(fx, rank, "cuda"), StepConfig(share_frontend=True, stack_passes=True), SP_MASK)
    loss.backward()
    sync.finish()
    grads = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).cpu()
    opt.step()
    weights = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu()
    gathered = [torch.empty_like(weights.cuda()) for _ in range(world)]
    dist.all_gather(gathered, weights.cuda())
    same_weights = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        q.put((grads.numpy(), bool(same_weights)))
    dist.barrier()
    dist.destroy_process_group()


def test_dp2_nccl_gradients_equal_single_gpu_accumulation():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, same_weights = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same_weights

    import onebit_b200 as ob
    from onebit_b200.training import StepConfig, cotraining_loss
    fx = dict(np.load(os.path.join(GOLDEN, "conformer_step.npz")))
    torch.manual_seed(3)
    model = ob.ConformerASR(**CFG).train().cuda()
    cfg = StepConfig(share_frontend=True, stack_passes=True)
    for r in range(2):                                      # gradient accumulation over the two micro-batches
        loss, _ = cotraining_loss(model, _shard(fx, r, "cuda"), cfg, SP_MASK)
        (0.5 * loss).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]).cpu().numpy()
    assert got.shape == ref.shape
    scale = float(np.abs(ref).max())
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-6 * scale)
