"""agnn_gather_reduce / agnn_rowscale_sum vs straight torch arithmetic on the CPU
(the oracle's mean_into_copy / sum_into formulations, oracle/intree.py)."""
import numpy as np
import pytest
import torch

from analysisgnn_b200 import _lib, graph, ops
from oracle import intree as oi
from tests.util import DEV, BF16_REL, FP32_REL, assert_close

pytestmark = pytest.mark.gpu


def _graph(n_rows, n_cols, e, seed, zipf=False):
    rng = np.random.default_rng(seed)
    if zipf:
        row = np.minimum(rng.zipf(1.3, e) - 1, n_rows - 1)
    else:
        row = rng.integers(0, n_rows, e)
    col = rng.integers(0, n_cols, e)
    return torch.as_tensor(np.stack((row, col)), dtype=torch.long)


@pytest.mark.parametrize("f", [4, 32, 64, 128, 256, 512])
@pytest.mark.parametrize("mean", [False, True])
def test_segment_reduce_forward_backward(f, mean):
    n_rows, n_cols, e = 301, 257, 2000
    ei = _graph(n_rows, n_cols, e, f)
    torch.manual_seed(f)
    src = torch.randn(n_cols, f)
    self_add = torch.randn(n_rows, f)
    g = torch.randn(n_rows, f)
    s1 = src.clone().requires_grad_(True)
    a1 = self_add.clone().requires_grad_(True)
    if mean:
        want = oi.mean_into_copy(s1[ei[1]], ei[0], a1)
    else:
        want = oi.sum_into(s1[ei[1]], ei[0], n_rows)
    want.backward(g)
    csr = graph.TypedCSR(ei.to(DEV), None, n_rows, n_cols=n_cols)
    s2 = src.to(DEV).requires_grad_(True)
    a2 = self_add.to(DEV).requires_grad_(True)
    got = ops.segment_mean_self(s2, a2, csr) if mean else ops.segment_sum(s2, csr)
    got.backward(g.to(DEV))
    assert_close(got, want, FP32_REL, "forward")
    assert_close(s2.grad, s1.grad, FP32_REL, "d src")
    if mean:
        assert_close(a2.grad, a1.grad, FP32_REL, "d self")


def test_transposed_view_reduces_the_other_way():
    n_rows, n_cols, e, f = 90, 140, 700, 64
    ei = _graph(n_rows, n_cols, e, 3)
    x = torch.randn(n_rows, f)
    want = oi.sum_into(x[ei[0]], ei[1], n_cols)
    csr = graph.TypedCSR(ei.to(DEV), None, n_rows, n_cols=n_cols)
    got = ops.segment_sum(x.to(DEV), csr.t())
    assert_close(got, want, FP32_REL)


def test_skewed_degrees_and_empty_rows():
    n, e, f = 5000, 60000, 256
    ei = _graph(n, n, e, 5, zipf=True)
    x = torch.randn(n, f)
    want = oi.mean_into_copy(x[ei[1]], ei[0], x)
    csr = graph.TypedCSR(ei.to(DEV), None, n)
    got = ops.segment_mean_self(x.to(DEV), x.to(DEV), csr)
    assert_close(got, want, FP32_REL)


def test_no_edges_at_all():
    n, f = 17, 32
    ei = torch.zeros((2, 0), dtype=torch.long)
    x = torch.randn(n, f)
    csr = graph.TypedCSR(ei.to(DEV), None, n)
    assert_close(ops.segment_sum(x.to(DEV), csr), torch.zeros(n, f) + 0 * x, FP32_REL)
    got = ops.segment_mean_self(x.to(DEV), x.to(DEV), csr)
    assert_close(got, x, FP32_REL)


def test_bf16_mode():
    n, e, f = 1000, 9000, 256
    ei = _graph(n, n, e, 6)
    x = torch.randn(n, f)
    want = oi.mean_into_copy(x[ei[1]], ei[0], x)
    csr = graph.TypedCSR(ei.to(DEV), None, n)
    xb = x.to(DEV, torch.bfloat16)
    got = ops.segment_mean_self(xb, xb, csr)
    assert got.dtype == torch.bfloat16
    assert_close(got.float(), want, BF16_REL)


def test_multi_relation_concat_and_sum_modes():
    """The relation-fused launch: every relation in its own output slice (concat) or summed."""
    n, f, r = 400, 64, 5
    rng = np.random.default_rng(11)
    e = 3000
    ei = torch.as_tensor(np.stack((rng.integers(0, n, e), rng.integers(0, n, e))), dtype=torch.long)
    et = torch.as_tensor(rng.integers(0, r - 1, e))               # the last relation stays empty
    h = torch.randn(n, r * f)
    x = torch.randn(n, f)
    csr = graph.TypedCSR(ei.to(DEV), et.to(DEV), n, n_rel=r)
    hd, xd = h.to(DEV), x.to(DEV)
    out = torch.empty((n, (r + 1) * f), device=DEV)
    rels = [ops.Rel(csr.fwd.rowptr[k], csr.fwd.col, hd[:, k * f:(k + 1) * f], out_col=(k + 1) * f,
                    flags=_lib.REL_IDENTITY_IF_EMPTY) for k in range(r)]
    ops.gather_reduce(rels, out, f, mean=True, concat=True, self_add=xd, copy=xd, copy_col=0)
    assert_close(out[:, :f], x, 0.0, "copy slice")
    for k in range(r):
        pick = et == k
        hk = h[:, k * f:(k + 1) * f]
        want = oi.mean_into_copy(hk[ei[1][pick]], ei[0][pick], x) if pick.any() else hk   # gnn.py:67-69
        assert_close(out[:, (k + 1) * f:(k + 2) * f], want, FP32_REL, f"relation {k}")
    tot = torch.empty((n, f), device=DEV)
    rels = [ops.Rel(csr.fwd.rowptr[k], csr.fwd.col, hd[:, k * f:(k + 1) * f]) for k in range(r)]
    ops.gather_reduce(rels, tot, f, mean=True, concat=False, self_add=xd)
    want = x.clone()
    for k in range(r):
        pick = et == k
        hk = h[:, k * f:(k + 1) * f]
        want = want + oi.mean_into_copy(hk[ei[1][pick]], ei[0][pick], torch.zeros(n, f))
    assert_close(tot, want, FP32_REL, "sum mode")


def test_linearity_at_full_size():
    """BASELINE config 1 size: gather(a + b) == gather(a) + gather(b), and the column sums of an
    un-normalised gather equal the degree-weighted column sums of the input."""
    from analysisgnn_b200 import synth
    b = synth.intree_batch(100, 500, 1, in_features=8, metrical=False)
    n, f = b["x"].shape[0], 256
    ei = b["edge_index"].to(DEV)
    csr = graph.TypedCSR(ei, None, n)
    torch.manual_seed(0)
    a, c = torch.randn(n, f, device=DEV), torch.randn(n, f, device=DEV)
    ga, gc, gac = ops.segment_sum(a, csr), ops.segment_sum(c, csr), ops.segment_sum(a + c, csr)
    assert_close(gac, ga + gc, FP32_REL)
    outdeg = torch.bincount(b["edge_index"][1], minlength=n).double().to(DEV)
    want = (a.double() * outdeg[:, None]).sum(0)
    assert_close(ga.double().sum(0), want, FP32_REL)


def test_tf32_pair_output_is_what_the_gemm_consumes():
    """out_lo: the gather writes rna_tf32(y) and rna_tf32(y - hi) -- bit-identical to agnn_split_tf32 of the plain result."""
    from analysisgnn_b200 import linalg
    n, e, f = 700, 5000, 256
    ei = _graph(n, n, e, 12)
    x = torch.randn(n, f, device=DEV)
    csr = graph.TypedCSR(ei.to(DEV), None, n)
    rel = [ops.Rel(csr.fwd.rowptr[0], csr.fwd.col, x, out_col=f)]
    plain = torch.empty((n, 2 * f), device=DEV)
    ops.gather_reduce(rel, plain, f, mean=True, concat=True, copy=x, copy_col=0)
    pair = torch.empty((2, n, 2 * f), device=DEV)
    ops.gather_reduce(rel, pair[0], f, mean=True, concat=True, copy=x, copy_col=0, out_lo=pair[1])
    want = linalg.split(plain)
    assert torch.equal(pair[0], want.hi) and torch.equal(pair[1], want.lo)


@pytest.mark.parametrize("hubs", [False, True])
def test_f16_pair_output_is_what_the_gemm_consumes(hubs):
    """agnn_gather_reduce_f16: out / out_lo receive fp16(s y) and fp16(s y - hi), s from the amax scalar --
    bit-identical to agnn_split_f16 of the plain result under the same scalar; also for hub rows (>= HEAVY_ROW entries),
    which the multi-warp kernels write."""
    from analysisgnn_b200 import linalg
    f = 256
    if hubs:
        rng = np.random.default_rng(5)
        n = 400
        row = np.concatenate([np.full(9000, 3), np.full(4096, 17), rng.integers(0, n, 5000)])
        rng.shuffle(row)
        ei = torch.as_tensor(np.stack((row, rng.integers(0, n, len(row)))), dtype=torch.long)
    else:
        n = 700
        ei = _graph(n, n, 5000, 12)
    torch.manual_seed(1)
    x = torch.randn(n, f, device=DEV) * 0.03
    csr = graph.TypedCSR(ei.to(DEV), None, n)
    if hubs:
        assert int(csr.fwd.n_heavy[0]) == 2
    rel = [ops.rel_of(csr.fwd, 0, x, out_col=f, n_edges=ei.shape[1])]
    plain = torch.empty((n, 2 * f), device=DEV)
    ops.gather_reduce(rel, plain, f, mean=True, concat=True, self_add=x, copy=x, copy_col=0)
    amax = linalg.amax_into(linalg.new_amax(DEV), x)       # |mean with the self term| <= 2 amax: inside the headroom
    pair = torch.empty((2, n, 2 * f), dtype=torch.float16, device=DEV)
    ops.gather_reduce(rel, pair[0], f, mean=True, concat=True, self_add=x, copy=x, copy_col=0, out_lo=pair[1],
                      pair_amax=amax)
    want = linalg.split_f16(plain, amax)
    assert torch.equal(pair[0], want.hi) and torch.equal(pair[1], want.lo)
    assert not torch.isinf(pair[0]).any()
    back = (pair[0].double() + pair[1].double()) / float(linalg.f16_scale(amax))
    assert_close(back, plain, 1e-6, "decoded pair")


@pytest.mark.parametrize("mean", [False, True])
def test_heavy_rows_are_split_across_warps(mean):
    """Hub rows (>= HEAVY_ROW entries) take the multi-warp path: chunk partials + ordered combine.  Same numbers as the
    plain formulation, deterministic from run to run, forward and backward."""
    rng = np.random.default_rng(3)
    n_rows, n_cols, f = 300, 5000, 64
    hr = _lib.HEAVY_ROW
    row = np.concatenate([np.full(20000, 7), np.full(hr, 11), np.full(hr - 80, 13), rng.integers(0, n_rows, 6000)])
    rng.shuffle(row)
    col = rng.integers(0, n_cols, len(row))
    ei = torch.as_tensor(np.stack((row, col)), dtype=torch.long)
    src = torch.randn(n_cols, f)
    base = torch.randn(n_rows, f)
    g = torch.randn(n_rows, f)
    s1 = src.clone().requires_grad_(True)
    want = oi.mean_into_copy(s1[ei[1]], ei[0], base) if mean else oi.sum_into(s1[ei[1]], ei[0], n_rows)
    want.backward(g)
    csr = graph.TypedCSR(ei.to(DEV), None, n_rows, n_cols=n_cols)
    want_heavy = np.flatnonzero(np.bincount(row, minlength=n_rows) >= _lib.HEAVY_ROW).tolist()
    assert 7 in want_heavy and 11 in want_heavy
    n_h, cap = int(csr.fwd.n_heavy[0]), csr.fwd.heavy_cap
    assert csr.fwd.heavy[0, :n_h].tolist() == want_heavy                     # ascending row order
    chunks = -(-np.bincount(row, minlength=n_rows)[want_heavy] // _lib.HEAVY_CHUNK)
    assert csr.fwd.heavy[0, cap:cap + n_h].tolist() == (np.cumsum(chunks) - chunks).tolist()
    s2 = src.to(DEV).requires_grad_(True)
    got = ops.segment_mean_self(s2, base.to(DEV), csr) if mean else ops.segment_sum(s2, csr)
    got.backward(g.to(DEV))
    assert_close(got, want, 2 * FP32_REL, "forward")        # 20 000-term sums: order differs from index_add_
    assert_close(s2.grad, s1.grad, FP32_REL, "d src")
    again = ops.segment_mean_self(s2, base.to(DEV), csr) if mean else ops.segment_sum(s2, csr)
    assert torch.equal(got, again)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mean", [False, True])
def test_two_heavy_relations_on_one_row_are_summed_in_order(dtype, mean):
    """COMBINE_SUM with hub rows in SEVERAL relations (the backward of a hub source node, ops._HeteroSageLayer): the
    row's heavy relations are added by ONE warp in relation order -- no atomics, no lost update, bit-identical from
    run to run -- and agree with the plain formulation."""
    rng = np.random.default_rng(11)
    n_rows, n_cols, f = 200, 3000, 128
    rows = [np.concatenate([np.full(9000, 5), np.full(5000, 9), rng.integers(0, n_rows, 3000)]),     # rel 0: 5, 9
            np.concatenate([np.full(4096, 5), rng.integers(0, n_rows, 2000)]),                       # rel 1: 5
            np.concatenate([np.full(6000, 9), np.full(4500, 5), rng.integers(0, n_rows, 1000)]),     # rel 2: 9, 5
            rng.integers(0, n_rows, 1500)]                                                           # rel 3: light
    eis = []
    for r in rows:
        rng.shuffle(r)
        eis.append(torch.as_tensor(np.stack((r, rng.integers(0, n_cols, len(r)))), dtype=torch.long))
    torch.manual_seed(2)
    srcs = [torch.randn(n_cols, f) for _ in rows]
    base = torch.randn(n_rows, f)
    want = base.to(dtype).double()
    for ei, s in zip(eis, srcs):
        sd = s.to(dtype).double()
        part = torch.zeros(n_rows, f, dtype=torch.float64).index_add_(0, ei[0], sd[ei[1]])
        if mean:
            part /= torch.bincount(ei[0], minlength=n_rows).clamp(min=1).double().unsqueeze(1)
        want += part
    csrs = [graph.TypedCSR(ei.to(DEV), None, n_rows, n_cols=n_cols) for ei in eis]
    assert [int(c.fwd.n_heavy[0]) for c in csrs] == [2, 1, 2, 0]
    rels = [ops.rel_of(c.fwd, 0, s.to(DEV, dtype), n_edges=ei.shape[1]) for c, s, ei in zip(csrs, srcs, eis)]
    outs = []
    for _ in range(3):
        out = torch.empty((n_rows, f), dtype=dtype, device=DEV)
        ops.gather_reduce(rels, out, f, mean=mean, concat=False, self_add=base.to(DEV, dtype))
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert_close(outs[0].double(), want, 2 * FP32_REL if dtype == torch.float32 else BF16_REL, "sum over relations")


@pytest.mark.parametrize("concat", [False, True])
def test_many_heavy_rows_zipf_degrees(concat):
    """A Zipf-like degree tail: a hundred rows of 500 .. 6000 entries per relation, three relations whose heavy rows
    partly coincide.  Chunks are found through the lists' chunk prefix (binary search) -- same numbers as the plain
    formulation, bit-identical from run to run."""
    rng = np.random.default_rng(5)
    n_rows, n_cols, f = 1500, 4000, 128
    eis = []
    for r in range(3):
        hubs = rng.choice(n_rows, 100, replace=False)
        degs = rng.integers(_lib.HEAVY_ROW - 12, 6000, len(hubs))
        row = np.concatenate([np.repeat(hubs, degs), rng.integers(0, n_rows, 8000)])
        rng.shuffle(row)
        eis.append(torch.as_tensor(np.stack((row, rng.integers(0, n_cols, len(row)))), dtype=torch.long))
    torch.manual_seed(4)
    srcs = [torch.randn(n_cols, f) for _ in eis]
    base = torch.randn(n_rows, f)
    parts = []
    for ei, s in zip(eis, srcs):
        part = torch.zeros(n_rows, f, dtype=torch.float64).index_add_(0, ei[0], s.double()[ei[1]])
        parts.append(part / torch.bincount(ei[0], minlength=n_rows).clamp(min=1).double().unsqueeze(1))
    want = torch.cat(parts, dim=1) if concat else base.double() + sum(parts)
    csrs = [graph.TypedCSR(ei.to(DEV), None, n_rows, n_cols=n_cols) for ei in eis]
    assert all(int(c.fwd.n_heavy[0]) >= 90 for c in csrs)
    rels = [ops.rel_of(c.fwd, 0, s.to(DEV), n_edges=ei.shape[1], out_col=(k * f if concat else 0))
            for k, (c, s, ei) in enumerate(zip(csrs, srcs, eis))]
    outs = []
    for _ in range(2):
        out = torch.empty((n_rows, (3 if concat else 1) * f), dtype=torch.float32, device=DEV)
        ops.gather_reduce(rels, out, f, mean=True, concat=concat, self_add=None if concat else base.to(DEV))
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    assert_close(outs[0].double(), want, 2 * FP32_REL, "zipf tail")


def test_degree_bound_hint_only_removes_the_hub_row_launches():
    """graph.set_degree_bound below the hub-row threshold: no hub-row lists, two launches fewer per aggregation, same
    values -- also when the hint is WRONG (the 700-entry row is then reduced by its own warp in edge order instead of
    chunk by chunk: equal up to fp32 summation order; every other row bit for bit)."""
    from analysisgnn_b200 import _lib, graph, ops
    torch.manual_seed(3)
    n, f = 600, 64
    src = torch.randint(0, n, (9000,), device=DEV)
    dst = torch.randint(1, n, (9000,), device=DEV)
    dst[:700] = 0                                            # a hub row: above HEAVY_ROW = 512
    ei = torch.stack([src, dst])
    x = torch.randn(n, f, device=DEV)
    outs, launches = [], []
    for bound in (None, 500):
        old = graph.set_degree_bound(bound)
        try:
            graph.clear_cache()
            csr = graph.typed_csr(ei, None, n, 1, reduce_row=1)
            assert (csr.fwd.heavy is None) == (bound is not None)
            l0 = _lib.launches()
            outs.append(ops.segment_sum(x, csr))
            launches.append(_lib.launches() - l0)
        finally:
            graph.set_degree_bound(old)
            graph.clear_cache()
    assert torch.equal(outs[0][1:], outs[1][1:])
    assert_close(outs[1][:1], outs[0][:1], FP32_REL, "hub row")
    assert launches[1] == launches[0] - 2 and launches[1] >= 1, launches


@pytest.mark.parametrize("dtype,f", [(torch.float32, 256), (torch.float32, 64), (torch.bfloat16, 256), (torch.float32, 100)])
@pytest.mark.parametrize("mean,concat,with_self", [(True, True, False), (False, True, True), (True, False, True),
                                                   (False, False, False)])
def test_low_degree_launch_equals_the_row_per_warp_kernel(dtype, f, mean, concat, with_self):
    """One relation with < 1.5 entries per row takes the four-rows-per-warp kernel (AGNN_REL_LOW_DEGREE, set by
    ops._pack when the edge count is known): same CSR order per row, so bit-identical to the general kernel; rows of
    0 - 7 entries, a row count that is not a multiple of 4, and a hub row that goes to the hub-row kernels."""
    from analysisgnn_b200 import graph, ops
    torch.manual_seed(11)
    n, n_src = 1003, 700
    deg = (torch.rand(n) < 0.4).long() + (torch.rand(n) < 0.1).long() * 2
    deg[5] = 7
    deg[17] = 600                                            # >= HEAVY_ROW
    dst = torch.repeat_interleave(torch.arange(n), deg)
    e = int(dst.numel())
    assert e < 1.5 * n
    ei = torch.stack([dst, torch.randint(0, n_src, (e,))]).to(DEV)
    csr = graph.TypedCSR(ei, None, n, n_cols=n_src)
    x = torch.randn(n_src, f, device=DEV).to(dtype)
    sa = torch.randn(n, f, device=DEV).to(dtype) if with_self else None
    outs = []
    for known in (e, None):                                  # None: no hint, the general kernel
        y = torch.full((n, f), float("nan"), device=DEV, dtype=dtype)
        rel = [ops.rel_of(csr.fwd, 0, x, n_edges=known)]
        if known is None:
            rel[0].heavy_rows = rel[0].n_heavy = None        # (the hub-row workspace needs the edge count)
        ops.gather_reduce(rel, y, f, mean=mean, concat=concat, self_add=sa)
        outs.append(y)
    keep = torch.ones(n, dtype=torch.bool, device=DEV)
    keep[17] = False                                         # chunked (hub-row path) vs edge order: not bit-identical
    assert torch.equal(outs[0][keep], outs[1][keep])
    assert_close(outs[0][17:18].float(), outs[1][17:18].float(), FP32_REL if dtype == torch.float32 else BF16_REL, "hub row")
