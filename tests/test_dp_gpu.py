"""Data-parallel MIL training over NCCL: two ranks, each with half of a batch of bags, must end up
with the same averaged gradients, running means and parameters as one process stepping on the whole batch.
Needs two GPUs (skipped otherwise); the same exchange is covered on CPU with gloo in tests/test_host_cpu.py."""

import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(dev):
    from oracle import mil_oracle
    from stamp_b200.mil import VisionTransformer

    sd = mil_oracle.init_state_dict(dim_input=64, dim_output=2, dim_model=128, n_heads=2, dim_feedforward=128, seed=21)
    m = VisionTransformer(dim_output=2, dim_input=64, dim_model=128, n_layers=2, n_heads=2, dim_feedforward=128,
                          dropout=0.0, use_alibi=True)
    m.load_state_dict(sd)
    for _, ff in m.transformer.layers:
        ff[3].p = 0.0
        ff[5].p = 0.0
    bags, coords = mil_oracle.synthetic_bag(300, 64, seed=3, batch=4, signal=True)
    y = torch.nn.functional.one_hot(torch.arange(4) % 2, 2).float()
    return m.to(dev).train(), bags.to(dev), coords.to(dev), y.to(dev)


def _worker(rank: int, world: int, port: int, q) -> None:
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from stamp_b200 import train as T

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model, bags, coords, y = _make(dev)
    opt = T.FusedAdamW(model.parameters(), lr=1e-3)
    sl = slice(rank * 2, rank * 2 + 2)
    loss = T.data_parallel_step(model, opt, (bags[sl], coords[sl], None, y[sl]), None)
    rm = model.transformer.layers[0][0].mhsa.attentions[0].scale_distance.running_mean
    q.put((rank, float(loss), (opt.flat_grad * 0.5).cpu(), opt.flat_param.cpu(), float(rm)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_step(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from stamp_b200 import train as T

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r[0]: r[1:] for r in (q.get(timeout=300) for _ in procs)}
    for p in procs:
        p.join(timeout=60)

    model, bags, coords, y = _make(cuda_device)
    opt = T.FusedAdamW(model.parameters(), lr=1e-3)
    loss = T.data_parallel_step(model, opt, (bags, coords, None, y), None)
    g_ref, p_ref = opt.flat_grad.cpu(), opt.flat_param.cpu()
    rm_ref = float(model.transformer.layers[0][0].mhsa.attentions[0].scale_distance.running_mean)

    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])      # replicas stay identical
    assert abs(0.5 * (res[0][0] + res[1][0]) - float(loss)) < 1e-3                      # mean of the local losses
    # the running mean is the mean over the local distances, averaged across ranks; bf16 activations differ by
    # the batch-dependent mean distance they are scaled with, hence a bf16-level tolerance on the gradients
    assert abs(res[0][3] - rm_ref) < 2e-3 * rm_ref
    rel = float((res[0][1] - g_ref).norm() / g_ref.norm())
    assert rel < 2e-2, rel
    moved = (p_ref - res[0][2]).abs().max()
    assert float(moved) < 2.1e-3          # both took one AdamW step of size <= lr from the same start
