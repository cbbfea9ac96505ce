"""csrc/peer.cu: the gradient exchange fused with AdamW over peer memory, SyncBatchNorm statistics over peer memory.

world = 1 runs on any single GPU (the exchanges degenerate to local copies; the kernels, barriers and epochs are the same
code).  world = 2 spawns two processes and needs two GPUs with P2P access (gpurun --gpus 2); it is skipped otherwise.
References: torch.optim.AdamW on the rank-averaged gradient (what DistributedDataParallel + AdamW compute, main.py:515-517,
engine.py:105-108) and fp64 BatchNorm1d over the concatenated batch of all ranks (nn.SyncBatchNorm, main.py:512-514)."""
import os
import socket
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.LayerNorm(64), torch.nn.ReLU(), torch.nn.Linear(64, 19),
                               torch.nn.Linear(19, 3, bias=False))


def _optimizer_case(rank, world, dev, clip):
    """3 steps; every rank feeds different data; returns max relative parameter difference to the reference."""
    import torch.distributed as dist
    from vdetr_b200 import parallel
    no_decay = lambda n, p: p.dim() == 1 or n.endswith("bias")          # noqa: E731
    a, b = _model(0).to(dev), _model(0).to(dev)
    ref = torch.optim.AdamW([{"params": [p for n, p in a.named_parameters() if no_decay(n, p)], "weight_decay": 0.0},
                             {"params": [p for n, p in a.named_parameters() if not no_decay(n, p)], "weight_decay": 0.1}], lr=7e-4)
    pg = parallel.PeerGroup(dev)
    opt = parallel.FlatAdamW(b.named_parameters(), lr=7e-4, weight_decay=0.1, no_decay=no_decay, peer=pg)
    opt.world_scale = 1.0 / world
    assert opt.n % 4 != 0 or True
    worst = 0.0
    for it in range(3):
        g = torch.Generator(device=dev).manual_seed(100 * it + rank)
        x = torch.randn(50, 37, device=dev, generator=g)
        # reference: gradients averaged over the ranks (NCCL all-reduce), clip, AdamW
        ref.zero_grad(set_to_none=True)
        a(x).square().sum().backward()
        if world > 1:
            for p in a.parameters():
                dist.all_reduce(p.grad)
                p.grad.div_(world)
        want_norm = torch.nn.utils.clip_grad_norm_(a.parameters(), clip) if clip > 0 else None
        ref.step()
        # ours
        opt.zero_grad()
        b(x).square().sum().backward()
        norm = opt.clip_grad_norm_(clip) if clip > 0 else None
        opt.step()
        torch.cuda.synchronize(dev)
        pg.check()
        if clip > 0:
            assert abs(float(norm) - float(want_norm)) <= 1e-4 * float(want_norm), (float(norm), float(want_norm))
        for (n, p), q in zip(a.named_parameters(), b.parameters()):
            worst = max(worst, (p - q).abs().max().item() / (p.abs().max().item() + 1e-3))
    # every rank holds the same parameters, bit for bit
    if world > 1:
        mine = opt.flat_p.clone()
        other = mine.clone()
        dist.broadcast(other, 0)
        assert torch.equal(mine, other)
    sd = opt.state_dict()
    assert sd["exp_avg"].numel() == opt.n
    opt.load_state_dict(sd)
    return worst


def _syncbn_case(rank, world, dev, groups):
    """Training BatchNorm+ReLU on different rows per rank with SyncBatchNorm on, against fp64 BatchNorm1d over all rows."""
    import torch.distributed as dist
    from vdetr_b200 import ops, parallel
    cols, rows = 256, 300 + 40 * rank                                   # unequal row counts per rank
    pg = parallel.PeerGroup(dev)
    pg.enable_sync_batchnorm()
    try:
        gen = torch.Generator(device=dev).manual_seed(7 + rank)
        bns = [torch.nn.BatchNorm1d(cols).to(dev).train() for _ in range(groups)]
        for i, bn in enumerate(bns):
            gw = torch.Generator(device=dev).manual_seed(50 + i)
            with torch.no_grad():
                bn.weight.copy_(torch.rand(cols, device=dev, generator=gw) + 0.5)
                bn.bias.copy_(torch.randn(cols, device=dev, generator=gw) * 0.3)
        x = (torch.randn(rows, groups * cols, device=dev, generator=gen) * 2 + 3.0 + rank).requires_grad_(True)
        dy = torch.randn(rows, groups * cols, device=dev, generator=gen)
        for rep in range(3):                                            # several exchanges: epoch parity, slot reuse
            y = ops.bn_relu_train(x, bns[0]) if groups == 1 else ops.bn_relu_train_group(x, bns, "cl")
            gx, = torch.autograd.grad(y, x, dy, retain_graph=True)
            gparams = torch.autograd.grad(y, [b.weight for b in bns] + [b.bias for b in bns], dy)
        torch.cuda.synchronize(dev)
        pg.check()
        # reference over the concatenated batch
        counts = [300 + 40 * r for r in range(world)]
        xs = [torch.empty(c, groups * cols, device=dev) for c in counts]
        dys = [torch.empty(c, groups * cols, device=dev) for c in counts]
        if world > 1:
            for r in range(world):
                xs[r] = x.detach().clone() if r == rank else xs[r]
                dys[r] = dy.clone() if r == rank else dys[r]
                dist.broadcast(xs[r], r)
                dist.broadcast(dys[r], r)
        else:
            xs, dys = [x.detach()], [dy]
        xa = torch.cat(xs).double().requires_grad_(True)
        dya = torch.cat(dys).double()
        lo = sum(counts[:rank])
        worst = 0.0
        for i, bn in enumerate(bns):
            rb = torch.nn.BatchNorm1d(cols).to(dev).double().train()
            with torch.no_grad():
                rb.weight.copy_(bn.weight.double()); rb.bias.copy_(bn.bias.double())
            sl = slice(i * cols, (i + 1) * cols)
            for rep in range(3):
                want = torch.relu(rb(xa[:, sl]))
            gxa, gwa, gba = torch.autograd.grad(want, (xa, rb.weight, rb.bias), dya[:, sl])
            err_y = (y[:, sl].double() - want[lo:lo + rows]).abs().max().item() / want.abs().max().item()
            err_dx = (gx[:, sl].double() - gxa[lo:lo + rows, sl]).abs().max().item() / gxa.abs().max().item()
            # parameter gradients are this rank's share; their sum over ranks is the full-batch gradient
            gw_l, gb_l = gparams[i].clone(), gparams[groups + i].clone()
            if world > 1:
                dist.all_reduce(gw_l); dist.all_reduce(gb_l)
            err_w = (gw_l.double() - gwa).abs().max().item() / gwa.abs().max().item()
            err_b = (gb_l.double() - gba).abs().max().item() / gba.abs().max().item()
            err_rm = (bn.running_mean.double() - rb.running_mean).abs().max().item()
            err_rv = (bn.running_var.double() - rb.running_var).abs().max().item()
            worst = max(worst, err_y / 2e-5, err_dx / 2e-4, err_w / 2e-4, err_b / 2e-4, err_rm / 1e-5, err_rv / 1e-4)
        return worst
    finally:
        pg.disable_sync_batchnorm()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        res = {"opt": _optimizer_case(rank, world, dev, 0.0), "opt_clip": _optimizer_case(rank, world, dev, 0.1),
               "bn1": _syncbn_case(rank, world, dev, 1), "bn5": _syncbn_case(rank, world, dev, 5)}
        q.put((rank, res))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "ERROR " + repr(e) + "\n" + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("clip", [0.0, 0.1])
def test_peer_adamw_world1_matches_torch_adamw(clip):
    assert _optimizer_case(0, 1, torch.device("cuda", 0), clip) <= 2e-6


@pytest.mark.parametrize("groups", [1, 5])
def test_peer_syncbn_world1_matches_batchnorm(groups):
    assert _syncbn_case(0, 1, torch.device("cuda", 0), groups) <= 1.0


def test_peer_barrier_is_graph_capturable():
    from vdetr_b200 import parallel
    pg = parallel.PeerGroup(torch.device("cuda", 0))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        pg.barrier(0)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        pg.barrier(0)
        pg.barrier(1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    pg.check()
    assert pg.epoch.tolist()[:2] == [4, 3]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_peer_world2_optimizer_and_syncbn():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, res in out:
        assert isinstance(res, dict), f"rank {rank}: {res}"
        assert res["opt"] <= 2e-6 and res["opt_clip"] <= 2e-6, res
        assert res["bn1"] <= 1.0 and res["bn5"] <= 1.0, res
