"""world_size-2 gloo test (CPU) of the N>1 path's host logic: scene sharding, DDP gradient averaging, max-over-ranks
timing.  The model is the CPU oracle decoder (the CUDA modules cannot run here); the plumbing under test is
vdetr_b200.parallel, which bench.py uses unchanged with the NCCL backend."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import recipe
    from oracle import decoder_torch as odt
    from vdetr_b200 import parallel
    torch.manual_seed(0)
    torch.set_num_threads(2)
    dec = odt.OracleDecoder(num_layers=1, num_queries=8, dropout=0.0, mlp_dropout=0.0).eval()   # eval: BN uses running stats
    ddp = parallel.wrap_data_parallel(dec, find_unused_parameters=True)
    c = recipe.decoder_case(5, 4, 24)                       # global batch of 4 scenes
    lo, hi = parallel.shard_range(4, rank, world)

    def run(model, sl):
        feat = torch.from_numpy(c["feat"][:, sl])
        out, _ = model(feat, torch.from_numpy(c["xyz"][sl]), [torch.from_numpy(c["mins"][sl]), torch.from_numpy(c["maxs"][sl])],
                       torch.from_numpy(c["center_normalized"][sl]), torch.from_numpy(c["size_normalized"][sl]))
        return odt.synthetic_loss(out)
    ddp.zero_grad()
    run(ddp, slice(lo, hi)).backward()
    got = {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None}
    # single-process reference: mean over the two shards of the shard gradients
    ref = odt.OracleDecoder(num_layers=1, num_queries=8, dropout=0.0, mlp_dropout=0.0).eval()
    ref.load_state_dict(dec.state_dict())
    acc = {}
    for r in range(world):
        a, b = parallel.shard_range(4, r, world)
        ref.zero_grad()
        run(ref, slice(a, b)).backward()
        for n, p in ref.named_parameters():
            if p.grad is not None:
                acc[n] = acc.get(n, 0) + p.grad / world
    err = max(float((got[n] - acc[n]).abs().max() / (acc[n].abs().max() + 1e-12)) for n in got)
    # the flat single-all-reduce path used by bench.py must give the same averaged gradients
    flat_model = odt.OracleDecoder(num_layers=1, num_queries=8, dropout=0.0, mlp_dropout=0.0).eval()
    flat_model.load_state_dict(dec.state_dict())
    for p_ in flat_model.pointcls_heads.parameters():
        p_.requires_grad_(False)
    sync = parallel.FlatGradAllReduce(flat_model.parameters())
    sync.zero_()
    run(flat_model, slice(lo, hi)).backward()
    sync.sync_()
    err2 = max(float((p_.grad - acc[n]).abs().max() / (acc[n].abs().max() + 1e-12))
               for n, p_ in flat_model.named_parameters() if p_.requires_grad and n in acc)
    err = max(err, err2)
    t = parallel.max_over_ranks(10.0 + rank)
    q.put((rank, err, t, (lo, hi)))
    dist.destroy_process_group()


def test_shard_range_is_a_partition():
    from vdetr_b200 import parallel
    for gb in (1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            parts = [parallel.shard_range(gb, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == gb
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


@pytest.mark.timeout(300)
def test_flat_shard_range_tiles_the_vector_on_float4_boundaries():
    """The shards of the peer-memory optimizer (parallel.FlatAdamW(peer=...), csrc/peer.cu) partition [0, n) for any n and
    world size, with float4-aligned interior boundaries."""
    from vdetr_b200.parallel import flat_shard_range
    for n in (0, 1, 3, 4, 5, 17, 1000, 11_634_563, 11_634_564):
        for world in (1, 2, 3, 8):
            edges = [flat_shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
                assert hi == lo2 and lo <= hi and hi % 4 == 0
            assert all(lo % 4 == 0 for lo, _ in edges)


def test_ddp_gloo_world2_gradients_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[3] for r in res) == [(0, 2), (2, 4)]
    assert all(r[1] < 1e-4 for r in res), res          # DDP average == mean of shard gradients
    assert all(abs(r[2] - 11.0) < 1e-6 for r in res)   # max over ranks
