"""Data-parallel step: gradient arena + allreduce (host logic, gloo world_size 2 on CPU) and the
fused clip + AdamW kernels vs torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW (GPU)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn as nn

from analysisgnn_b200 import _lib
from analysisgnn_b200.train import DataParallelTrainer, GradArena, shard_indices
from tests.util import DEV, assert_close


def _toy(seed=0):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(7, 13), nn.ReLU(), nn.Linear(13, 5), nn.LayerNorm(5), nn.Linear(5, 3, bias=False))


def test_arena_views_and_zero_filled_unused_parameters():
    net = _toy()
    extra = nn.Linear(4, 4)                                  # never used in the loss
    params = list(net.parameters()) + list(extra.parameters())
    arena = GradArena(params, 4096)
    assert arena.numel % 4 == 0 and all(o % 4 == 0 for o in arena.offsets)
    net(torch.randn(9, 7)).sum().backward()
    arena.check_views()
    ref = _toy()
    ref(torch.randn(9, 7))                                    # different input: just shapes
    for p, o in zip(arena.params, arena.offsets):
        assert torch.equal(arena.grad[o:o + p.numel()].view_as(p), p.grad)
    assert float(extra.weight.grad.abs().sum()) == 0.0        # zeros, not None
    arena.zero_()
    assert float(arena.grad.abs().sum()) == 0.0 and float(net[0].weight.grad.abs().sum()) == 0.0
    net.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError):
        arena.check_views()


def test_shard_indices_partition_the_batch():
    for world in (1, 2, 4, 8):
        parts = [shard_indices(100, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(100))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_step_refuses_cpu_parameters():
    net = _toy()
    tr = DataParallelTrainer(net, world_size=1)
    net(torch.randn(3, 7)).sum().backward()
    with pytest.raises(_lib.AgnnError):
        tr.step()


def test_collect_mode_fills_the_same_arena():
    """collect_grads: detached .grad, one multi-tensor copy after backward, zeros for unused parameters."""
    a, b = _toy(3), _toy(3)
    ta = DataParallelTrainer(a, world_size=1)
    tb = DataParallelTrainer(b, world_size=1, collect_grads=True)
    torch.manual_seed(5)
    x = torch.randn(9, 7)
    for it in range(2):
        ta.zero_grad(), tb.zero_grad()
        assert all(p.grad is None for p in b.parameters())
        if it == 0:
            a(x).sum().backward(), b(x).sum().backward()
        else:                                                # only the first layer gets a gradient
            a[0](x).sum().backward(), b[0](x).sum().backward()
        tb.collect()
        assert torch.equal(ta.arena.grad, tb.arena.grad)
        tb.arena.check_views()
    tb.collect()                                             # idempotent
    b[0](x).sum().backward()                                 # accumulating on top of collected views still works
    a[0](x).sum().backward()
    assert torch.equal(ta.arena.grad, tb.arena.grad)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _toy(rank)                                         # every rank seeds differently ...
    net.add_module("bn", nn.BatchNorm1d(3))
    with torch.no_grad():
        net.bn.running_mean.fill_(float(rank + 1))
    tr = DataParallelTrainer(net)                            # ... and starts from rank 0's parameters and buffers
    assert tr.world_size == world
    ref = _toy(0)
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.equal(p, q)
    assert float(net.bn.running_mean[0]) == 1.0
    del net.bn
    torch.manual_seed(100)
    data = torch.randn(12, 7)
    tr.zero_grad()
    net(data[shard_indices(12, rank, world)]).sum().backward()
    tr.allreduce()
    out[rank] = tr.arena.grad.clone()
    dist.destroy_process_group()


def test_gloo_world2_allreduce_equals_full_batch_gradient():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        g0, g1 = out[0], out[1]
    assert torch.equal(g0, g1)
    net = _toy(0)
    tr = DataParallelTrainer(net, world_size=1)
    torch.manual_seed(100)
    net(torch.randn(12, 7)).sum().backward()
    n = tr.arena.numel                                       # (the workers' arenas end with the BatchNorm probe)
    assert_close(g0[:n], tr.arena.grad, 1e-6, "sum of shard gradients")


@pytest.mark.gpu
@pytest.mark.parametrize("max_norm", [1.0, 0.0, 1e6])
def test_fused_clip_adamw_matches_torch(max_norm):
    """Same gradients into both optimizers (copied from the arena), four steps."""
    torch.manual_seed(0)
    mk = lambda: nn.Sequential(nn.Linear(33, 257), nn.ReLU(), nn.Linear(257, 129), nn.LayerNorm(129),
                               nn.Linear(129, 5001, bias=False))
    a = mk().to(DEV)
    b = mk().to(DEV)
    b.load_state_dict(a.state_dict())
    tr = DataParallelTrainer(a, lr=5e-3, weight_decay=5e-3, max_norm=max_norm, world_size=1)
    opt = torch.optim.AdamW(b.parameters(), lr=5e-3, weight_decay=5e-3)
    for it in range(4):
        x = torch.randn(64, 33, device=DEV) * (10.0 if it % 2 else 0.1)
        tr.zero_grad()
        a(x).square().mean().backward()
        for p, q in zip(a.parameters(), b.parameters()):
            q.grad = p.grad.clone()
        tr.step()
        if max_norm:
            norm = torch.nn.utils.clip_grad_norm_(b.parameters(), max_norm)
            assert_close(tr.grad_norm[0], norm, 1e-5, "gradient norm")
        opt.step()
        for (n, p), q in zip(a.named_parameters(), b.parameters()):
            assert_close(p, q, 2e-6, f"step {it} {n}")


@pytest.mark.gpu
def test_grad_scale_is_the_world_average():
    torch.manual_seed(1)
    a = nn.Linear(16, 16).to(DEV)
    b = nn.Linear(16, 16).to(DEV)
    b.load_state_dict(a.state_dict())
    tr = DataParallelTrainer(a, max_norm=0.0, world_size=1)
    tr.world_size = 4                                        # pretend the arena holds a 4-rank sum
    opt = torch.optim.AdamW(b.parameters(), lr=5e-3, weight_decay=5e-3)
    x = torch.randn(8, 16, device=DEV)
    tr.zero_grad()
    (a(x).sum() * 4).backward()
    tr.allreduce = lambda: None
    tr.step()
    b(x).sum().backward()
    opt.step()
    assert_close(a.weight, b.weight, 2e-6)


@pytest.mark.gpu
def test_collect_mode_trains_identically():
    torch.manual_seed(2)
    mk = lambda: nn.Sequential(nn.Linear(20, 64), nn.ReLU(), nn.Linear(64, 9))
    a, b = mk().to(DEV), mk().to(DEV)
    b.load_state_dict(a.state_dict())
    ta = DataParallelTrainer(a, world_size=1)
    tb = DataParallelTrainer(b, world_size=1, collect_grads=True)
    for it in range(3):
        x = torch.randn(32, 20, device=DEV)
        for net, tr in ((a, ta), (b, tb)):
            tr.zero_grad()
            net(x).square().mean().backward()
            tr.step()
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.equal(p, q)


@pytest.mark.gpu
def test_graphed_step_replays_the_eager_step():
    """train.GraphedStep: CSR build + forward + backward (+ gradient collection) captured once and replayed
    must train exactly like the same step launched eagerly -- also after new data is written into the static
    input buffers (the CSR is rebuilt inside the graph)."""
    import bench
    from analysisgnn_b200 import graph, linalg, synth
    from analysisgnn_b200 import nn as ann
    from analysisgnn_b200.train import GraphedStep
    tasks = {"cadence": 4, "localkey": 50}
    b = synth.hetero_batch(3, 60, 11, task_dict=tasks)
    first = bench.batch_tensors(b)
    gen = torch.Generator().manual_seed(5)
    second = {}
    for k, v in first.items():                               # same shapes: other features / labels, edges reordered
        if k.startswith("x."):
            second[k] = torch.randn(v.shape, generator=gen)
        elif k.startswith("ei."):
            second[k] = v[:, torch.randperm(v.shape[1], generator=gen)]
        elif k.startswith("label."):
            second[k] = v[torch.randperm(v.shape[0], generator=gen)]
        else:
            second[k] = v.clone()

    def make():
        torch.manual_seed(0)
        net = ann.AnalysisEncoder(b["metadata"], 25, 32, 16, tasks, 2, dropout=0.0).to(DEV)
        net.train()
        return net, DataParallelTrainer(net, lr=1e-3, weight_decay=1e-2, max_norm=1.0, world_size=1, collect_grads=True)

    def step_fn(net, tr):
        def fwd_bwd(t):                                      # no clear_cache / begin_step here: GraphedStep's job
            d = bench.unflatten(t, b)
            tr.zero_grad()
            logits = net(d["pitch_spelling"], d["key_signature"], d["x_dict"], d["edge_index_dict"], d["batch_dict"],
                         d["batch_size"], None, None)
            loss = ann.multitask_ce(logits, d["labels"])
            loss.backward()
            tr.collect()
            return loss
        return fwd_bwd

    data = [{k: v.to(DEV) for k, v in x.items()} for x in (first, second)]
    net_e, tr_e = make()
    eager = step_fn(net_e, tr_e)
    net_g, tr_g = make()
    static = {k: v.clone() for k, v in data[0].items()}
    start = {k: v.detach().clone() for k, v in net_g.state_dict().items()}
    graphed = GraphedStep(step_fn(net_g, tr_g), static, warmup=2)      # warm-up steps do not call the optimizer
    net_g.load_state_dict(start)
    losses_e, losses_g = [], []
    for it in range(4):
        batch = data[it % 2]
        losses_e.append(float(eager(batch).detach()))
        tr_e.step()
        for k, v in batch.items():
            static[k].copy_(v)
        losses_g.append(float(graphed().detach()))
        tr_g.step()
    assert losses_e == losses_g, (losses_e, losses_g)
    for (n, p), q in zip(net_e.named_parameters(), net_g.parameters()):
        assert torch.equal(p, q), n
    assert graphed.launches_per_replay > 50
    # the learning rate is a device scalar: a scheduler's new value reaches the (replayable) optimizer launch
    tr_g.set_lr(0.0)
    before = [p.detach().clone() for p in net_g.parameters()]
    graphed()
    tr_g.weight_decay, keep = 0.0, tr_g.weight_decay
    tr_g.step()
    tr_g.weight_decay = keep
    for p, q in zip(net_g.parameters(), before):
        assert torch.equal(p, q)


def test_loader_shards_a_global_batch_across_ranks():
    """Host side of the data-parallel loader (sampler.ScoreGraphLoader.batch_ids): for every world size the ranks'
    shares of a global batch are disjoint, together they are the batch, an epoch visits every score once, and the
    order depends on (seed, epoch) only."""
    import torch
    from analysisgnn_b200 import sampler
    n_scores = 37
    node_ptr = [0]
    for s in range(n_scores):
        node_ptr.append(node_ptr[-1] + 5 + s % 4)
    corpus = sampler.Corpus(torch.zeros(node_ptr[-1], 2), torch.zeros((3, 0), dtype=torch.long), node_ptr)
    single = sampler.ScoreGraphLoader(corpus, subgraph_size=4, batch_size=8, seed=11)
    assert len(single) == 5
    for epoch in (0, 1):
        seen = []
        for index in range(len(single)):
            whole = single.batch_ids(epoch, index)
            seen += whole
            for world in (2, 3, 8):
                parts = [sampler.ScoreGraphLoader(corpus, 4, 8, seed=11, rank=r, world_size=world).batch_ids(epoch, index)
                         for r in range(world)]
                flat = [g for p in parts for g in p]
                if len(whole) >= world:
                    assert sorted(flat) == sorted(whole) and len(set(flat)) == len(flat)
                else:       # short last batch: padded by wrapping around the epoch order, every rank gets one score
                    assert set(whole) <= set(flat) and len(flat) == world and all(len(p) == 1 for p in parts)
                    assert sorted(set(flat) - set(whole)) == sorted(single.order(epoch)[:world - len(whole)])
                assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
                assert all(parts)                            # same number of steps on every rank (no allreduce hang)
        assert sorted(seen) == list(range(n_scores))
    for world in (4, 8):                                     # corpus size not divisible by batch_size * world
        lens = {len(sampler.ScoreGraphLoader(corpus, 4, 8, seed=11, rank=r, world_size=world)) for r in range(world)}
        assert lens == {5}
    assert single.order(0) != single.order(1)
    assert single.order(0) == sampler.ScoreGraphLoader(corpus, 4, 8, seed=11).order(0)
    assert single.order(0) != sampler.ScoreGraphLoader(corpus, 4, 8, seed=12).order(0)
    assert sampler.ScoreGraphLoader(corpus, 4, 8, seed=11, shuffle=False).order(3) == list(range(n_scores))


def test_hetero_batch_serves_the_reference_training_step():
    """sampler.HeteroBatch: the accesses of ``ContinualAnalysisGNN.common_step`` / ``create_mask_dict``
    (analysisgnn/models/analysis.py:926-968) on a batch dict of the shape ``ScoreGraphLoader.batch`` returns."""
    import torch
    from analysisgnn_b200 import sampler
    n, extra = 12, 5
    task_dict = {"cadence": 4, "localkey": 50}
    out = {
        "batch_size": n, "graph_ids": [3, 1],
        "x_dict": {"note": torch.randn(n + extra, 6)},
        "edge_index_dict": {("note", "onset", "note"): torch.zeros((2, 0), dtype=torch.long)},
        "batch_dict": {"note": torch.cat((torch.zeros(8, dtype=torch.long), torch.ones(n + extra - 8, dtype=torch.long)))},
        "num_sampled_nodes_dict": {"note": [n, extra]}, "num_sampled_edges_dict": {("note", "onset", "note"): [0]},
        "extras": {"pitch_spelling": torch.randint(0, 35, (n + extra,)), "key_signature": torch.randint(0, 15, (n + extra,)),
                   "cadence": torch.randint(0, 6, (n + extra,)), "localkey": torch.randint(0, 50, (n + extra,)),
                   "valid_label": torch.ones(n + extra, dtype=torch.long)},
    }
    batch = sampler.HeteroBatch(out)
    # analysis.py:948-961, verbatim access pattern
    x_dict = batch.x_dict
    batch_size = batch["note"].batch_size
    labels_dict = {k: batch["note"][k][:batch_size] for k in task_dict.keys() if k in batch["note"].keys()}
    pitch_spelling = batch["note"].pitch_spelling
    key_signature = batch["note"].key_signature
    labels_dict = {k: torch.where(labels_dict[k] < task_dict[k], labels_dict[k], torch.zeros_like(labels_dict[k]))
                   for k in labels_dict.keys()}
    assert batch.edge_index_dict is out["edge_index_dict"] and batch.batch_dict is out["batch_dict"]
    assert batch.num_sampled_edges_dict is out["num_sampled_edges_dict"]
    assert batch.num_sampled_nodes_dict is out["num_sampled_nodes_dict"]
    assert "valid_label" in batch["note"].keys() and "has_cadence" not in batch["note"].keys()
    valid = batch["note"]["valid_label"][:batch_size].bool()
    assert batch_size == n and set(labels_dict) == set(task_dict) and int(labels_dict["cadence"].max()) < 4
    assert x_dict["note"].shape[0] == n + extra and pitch_spelling.shape == key_signature.shape == (n + extra,)
    assert valid.all() and batch.node_types == ["note"] and batch["note"].x is x_dict["note"]
    with pytest.raises(AttributeError):
        batch["note"].not_there


def test_loader_subgraph_sample_ratio_visits():
    """``subgraph_sample_ratio`` (the reference passes 0.5 to every loader, datamodules/analysis.py:276): a score of n
    notes is visited max(1, ceil(ratio n / subgraph_size)) times per epoch; the visits of one score land in different
    batches and -- should two meet -- draw different windows; ranks split every global batch disjointly whatever the
    world size, with the same windows; ``None`` leaves the one-visit epoch exactly as it was."""
    import torch
    from analysisgnn_b200 import sampler
    from oracle import graph as og
    sizes = [40, 100, 101, 199, 200, 201, 450, 1000, 77, 12]
    node_ptr = [0]
    for n in sizes:
        node_ptr.append(node_ptr[-1] + n)
    corpus = sampler.Corpus(torch.zeros(node_ptr[-1], 2), torch.zeros((3, 0), dtype=torch.long), node_ptr)
    plain = sampler.ScoreGraphLoader(corpus, subgraph_size=100, batch_size=4, seed=5)
    assert plain.visits == [1] * len(sizes) and len(plain) == 3
    assert sorted(plain.order(0)) == list(range(len(sizes)))
    assert [plain.batch_ids(0, i) for i in range(3)] == [plain.order(0)[4 * i:4 * i + 4] for i in range(3)]
    assert plain.batch_draws(0, 0) == [0, 0, 0, 0]

    half = sampler.ScoreGraphLoader(corpus, subgraph_size=100, batch_size=4, seed=5, subgraph_sample_ratio=0.5)
    assert half.visits == [1, 1, 1, 1, 1, 2, 3, 5, 1, 1]            # ceil(n / 200), at least 1
    two = sampler.ScoreGraphLoader(corpus, subgraph_size=100, batch_size=4, seed=5, subgraph_sample_ratio=2)
    assert two.visits == [1, 2, 3, 4, 4, 5, 9, 20, 2, 1]            # ceil(n / 50)
    for loader in (half, two):
        n_batches = len(loader)
        assert n_batches == -(-sum(loader.visits) // 4)
        for epoch in range(2):
            order = loader.order(epoch)
            assert sorted(order) == sorted(g for g, v in enumerate(loader.visits) for _ in range(v))
            got = [loader.batch_ids(epoch, i) for i in range(n_batches)]
            assert [g for b in got for g in b] == order
            assert max(map(len, got)) - min(map(len, got)) <= 1
            for b in got:                                           # no score twice while visits <= batches
                dup = [g for g in set(b) if b.count(g) > 1]
                assert all(loader.visits[g] > n_batches for g in dup), (b, dup)
        assert loader.order(0) != loader.order(1)
    # the 1000-note score has 20 visits in 13 batches of ``two``: some batch holds it twice, with different draws
    met = [i for i in range(len(two)) if two.batch_ids(0, i).count(7) > 1]
    assert met
    ids, draws, starts = two.batch_ids(0, met[0]), two.batch_draws(0, met[0]), two.window_starts(0, met[0])
    assert sorted(d for g, d in zip(ids, draws) if g == 7) == list(range(ids.count(7)))
    step_seed = sampler.rng_u64(two.seed, 0x424154, 0, met[0], 0)
    assert starts == [og.window_start(step_seed, g, sizes[g], 100, d) for g, d in zip(ids, draws)]
    assert len({s for g, s in zip(ids, starts) if g == 7}) > 1
    assert all(0 <= s <= max(sizes[g] - 100, 0) for g, s in zip(ids, starts))
    # data parallel: disjoint shares, the same windows as the single-rank loader
    for world in (2, 3):
        for index in range(len(two)):
            parts = [sampler.ScoreGraphLoader(corpus, 100, 4, seed=5, rank=r, world_size=world, subgraph_sample_ratio=2)
                     for r in range(world)]
            whole = list(zip(two.batch_ids(0, index), two.window_starts(0, index)))
            shares = [list(zip(p.batch_ids(0, index), p.window_starts(0, index))) for p in parts]
            assert all(shares), "every rank steps"
            if len(whole) >= world:
                assert sorted(x for s in shares for x in s) == sorted(whole)
                assert all(share == whole[r::world] for r, share in enumerate(shares))
    with pytest.raises(ValueError):
        sampler.ScoreGraphLoader(corpus, 100, 4, subgraph_sample_ratio=0.0)
