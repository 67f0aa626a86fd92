"""Parity of the CUDA LightGCN path (through the C ABI) against the oracle and the golden vectors."""
import numpy as np
import pytest
import torch

import oracle as O
from conftest import (assert_close, assert_topk_matches_scores, filter_dict, golden_files, load_golden)

pytestmark = pytest.mark.gpu
LG = golden_files("lightgcn")


def _model_from_golden(g):
    from hnm_recommendation_b200 import LightGCN
    alpha = None if np.isnan(g["alpha"]) else float(g["alpha"])
    m = LightGCN(int(g["num_users"]), int(g["num_items"]), embedding_dim=int(g["embedding_dim"]),
                 num_layers=int(g["num_layers"]), top_k=int(g["top_k"]), alpha=alpha)
    m.load_state_dict({"embeddings.weight": torch.from_numpy(g["weight"])})
    m = m.to("cuda")
    ew = torch.from_numpy(g["edge_weight"]) if g["edge_weight"].size else None
    m.set_graph(torch.from_numpy(g["edge_index"]), ew)
    return m


@pytest.mark.parametrize("path", LG, ids=lambda p: p.split("lightgcn_")[-1][:-4])
def test_golden_forward_and_scores(hnm_lib, path):
    g = load_golden(path)
    m = _model_from_golden(g)
    ue, ie = m.forward()
    assert ue.shape == (m.num_users, m.embedding_dim) and ie.shape == (m.num_items, m.embedding_dim)
    assert_close(ue, g["user_emb"], what="user_emb")          # P1: rtol 1e-5 (BASELINE.json north_star)
    assert_close(ie, g["item_emb"], what="item_emb")
    uids = torch.from_numpy(g["user_ids"])
    assert_close(m.predict_all_items(uids), g["scores"], what="scores")
    assert_close(m.predict(uids, torch.from_numpy(g["item_ids"])), g["pair_scores"], what="pair")


@pytest.mark.parametrize("path", LG, ids=lambda p: p.split("lightgcn_")[-1][:-4])
def test_golden_recommend(hnm_lib, path):
    g = load_golden(path)
    m = _model_from_golden(g)
    k = int(g["top_k"])
    uids = torch.from_numpy(g["user_ids"])
    # end to end (own embeddings): equal to the reference's list up to near-ties (P3)
    got = m.recommend(uids)
    assert got.dtype == torch.int64 and tuple(got.shape) == (uids.numel(), k)
    s64 = O.exact_scores_fp64(torch.from_numpy(g["user_emb"]), torch.from_numpy(g["item_emb"]), uids)
    assert_topk_matches_scores(got, s64, k)
    fd = filter_dict(g)
    got_f = m.recommend(uids, filter_items=fd)
    s64f = O.apply_filter(s64.clone(), uids, fd)
    assert_topk_matches_scores(got_f, s64f, k)


@pytest.mark.parametrize("path", LG, ids=lambda p: p.split("lightgcn_")[-1][:-4])
def test_golden_topk_stagewise_bit_exact(hnm_lib, path):
    """P2: fed with the reference's own embedding tensors, ids and fp64 scores are bit-identical to the oracle."""
    from hnm_recommendation_b200 import engine
    g = load_golden(path)
    k = int(g["top_k"])
    ue, ie = torch.from_numpy(g["user_emb"]), torch.from_numpy(g["item_emb"])
    uids = torch.from_numpy(g["user_ids"])
    want_ids, want_s = O.recommend_exact(ue, ie, uids, k)
    ids, sc = engine.topk_exact(ue.cuda(), ie.cuda(), uids, k)
    assert torch.equal(ids.cpu(), want_ids)
    assert torch.equal(sc.cpu(), want_s)
    fd = filter_dict(g)
    want_ids, want_s = O.recommend_exact(ue, ie, uids, k, fd)
    ids, sc = engine.topk_exact(ue.cuda(), ie.cuda(), uids, k, engine.exclusion_csr(uids, fd, "cuda"))
    assert torch.equal(ids.cpu(), want_ids) and torch.equal(sc.cpu(), want_s)
    # and they agree with the reference's fp32 list wherever that list is not a near-tie
    assert_topk_matches_scores(torch.from_numpy(g["topk_canonical"]), O.exact_scores_fp64(ue, ie, uids), k)


def _config1():
    from hnm_recommendation_b200 import synth
    data = synth.interactions(*synth.CONFIG1, seed=42)
    n = data.num_users + data.num_items
    return data, synth.xavier_embeddings(n, 64, seed=42)


def test_config1_forward_matches_oracle(hnm_lib):
    """BASELINE.json configs[0]: 10k users x 5k items x 200k interactions, d=64, L=3."""
    from hnm_recommendation_b200 import LightGCN
    data, w = _config1()
    ei = data.edge_index()
    orc = O.LightGCNOracle(data.num_users, data.num_items, 64, 3, 12, weight=w)
    orc.set_graph(ei)
    ou, oi = orc.forward()
    m = LightGCN(data.num_users, data.num_items).to("cuda")
    m.load_state_dict({"embeddings.weight": w})
    m.set_graph(ei)
    assert int(m.graph.nnz) == 2 * 200_000 + 15_000
    assert torch.equal(m.graph.rowptr.cpu().long(), orc.graph[0])
    assert torch.equal(m.graph.col.cpu().long(), orc.graph[1])
    assert_close(m.graph.dis, orc.graph[3], rtol=2e-7, atol_scale=0, what="dis")
    gu, gi = m.forward()                                   # train mode + autograd on: differentiable, as the reference's
    assert gu.requires_grad and gu.grad_fn is not None
    assert_close(gu.detach(), ou, what="users")
    assert_close(gi.detach(), oi, what="items")
    assert m.forward()[0].data_ptr() != gu.data_ptr()      # ... and recomputed on every call, like the reference
    m.eval()
    gu, gi = m.forward()
    assert not gu.requires_grad
    assert_close(gu, ou, what="users (eval)")
    # eval mode: the cached forward returns the same buffer until the weights change
    assert m.forward()[0].data_ptr() == gu.data_ptr()
    with torch.no_grad():
        m.embeddings.weight.mul_(2.0)
    gu2, _ = m.forward()
    assert_close(gu2, 2 * ou, what="users after in-place weight update")
    # a write through .data is invisible to the version counter: invalidate(), or the verify mode
    m.embeddings.weight.data.mul_(0.5)
    assert m.forward()[0].data_ptr() == gu2.data_ptr()     # stale by design of the key ...
    m.invalidate()
    assert_close(m.forward()[0], ou, what="users after invalidate()")
    m.cache_embeddings = "verify"
    m.invalidate()
    a = m.forward()[0]
    assert m.forward()[0].data_ptr() == a.data_ptr()
    m.embeddings.weight.data.mul_(3.0)
    assert_close(m.forward()[0], 3 * ou, what="users after a .data write in verify mode")
    m.load_state_dict({"embeddings.weight": w})            # load_state_dict invalidates by itself
    m.cache_embeddings = True
    assert_close(m.forward()[0], ou, what="users after load_state_dict")


def test_config1_topk_all_users_bit_exact(hnm_lib):
    from hnm_recommendation_b200 import engine
    data, w = _config1()
    orc = O.LightGCNOracle(data.num_users, data.num_items, 64, 3, 12, weight=w)
    orc.set_graph(data.edge_index())
    ou, oi = orc.forward()
    uids = torch.arange(data.num_users)
    want_ids, want_s = O.recommend_exact(ou, oi, uids, 12)
    ids, sc = engine.topk_exact(ou.cuda(), oi.cuda(), None, 12)
    assert torch.equal(ids.cpu(), want_ids) and torch.equal(sc.cpu(), want_s)
    # reference-faithful fp32 sgemm + stable sort agrees except at near-ties; count them
    ref32 = O.recommend(ou, oi, uids, 12)
    n_diff = int((ref32 != want_ids).any(dim=1).sum())
    assert n_diff <= 20, f"{n_diff} users differ between fp32-sgemm and exact ordering"
    # item-sharded: two shards merged == unsharded
    half = data.num_items // 2
    a = engine.topk_exact(ou.cuda(), oi[:half].cuda().contiguous(), None, 12, item_begin=0)
    b = engine.topk_exact(ou.cuda(), oi[half:].cuda().contiguous(), None, 12, item_begin=half)
    mi, ms = engine.merge_topk(torch.stack([a[0], b[0]]), torch.stack([a[1], b[1]]))
    assert torch.equal(mi.cpu(), want_ids) and torch.equal(ms.cpu(), want_s)


def test_merge_many_short_lists_with_sentinels(hnm_lib):
    """hnm_merge_topk: 24 item shards of 7 items each, k = 12 > shard size -- every list is the shard's exact
    top-7 followed by (-inf, INT64_MAX) sentinels (what ShardedLightGCN's item mode hands in for a shard smaller
    than k); the merge equals the exact top-12 of the whole catalog, and more than 64 lists are refused."""
    from hnm_recommendation_b200 import engine
    from hnm_recommendation_b200._lib import HnmError
    g = torch.Generator().manual_seed(11)
    U, shards, per, k, d = 300, 24, 7, 12, 32
    ue = torch.randn(U, d, generator=g)
    ie = torch.randn(shards * per, d, generator=g)
    ie[5] = ie[100]                                              # an exact tie across two shards: id ascending
    want_ids, want_s = O.recommend_exact(ue, ie, torch.arange(U), k)
    ids = torch.full((shards, U, k), torch.iinfo(torch.int64).max, dtype=torch.int64, device="cuda")
    sc = torch.full((shards, U, k), float("-inf"), dtype=torch.float64, device="cuda")
    for s in range(shards):
        a, b = engine.topk_exact(ue.cuda(), ie[s * per:(s + 1) * per].cuda().contiguous(), None, per, item_begin=s * per)
        ids[s, :, :per], sc[s, :, :per] = a, b
    mi, ms = engine.merge_topk(ids, sc)
    assert torch.equal(mi.cpu(), want_ids) and torch.equal(ms.cpu(), want_s)
    with pytest.raises(HnmError):
        engine.merge_topk(ids.repeat(3, 1, 1), sc.repeat(3, 1, 1))          # 72 lists


def test_ties_and_filter_edge_cases(hnm_lib):
    from hnm_recommendation_b200 import engine
    # identical item rows -> exact ties -> id ascending
    ue = torch.tensor([[1.0, 0.0, 2.0, 0.5]] * 3)
    ie = torch.tensor([[1.0, 1.0, 1.0, 1.0]] * 6 + [[2.0, 2.0, 2.0, 2.0]] * 2 + [[-1.0, 0, 0, 0]] * 2)
    ids, sc = engine.topk_exact(ue.cuda(), ie.cuda(), None, 5)
    assert ids.cpu().tolist() == [[6, 7, 0, 1, 2]] * 3
    # k == num_items, all items returned once
    ids, _ = engine.topk_exact(ue.cuda(), ie.cuda(), None, 10)
    assert sorted(ids[0].cpu().tolist()) == list(range(10))
    # filtering leaves fewer than k finite scores: -inf tail, smallest excluded ids first
    uids = torch.tensor([0, 1])
    excl = engine.exclusion_csr(uids, {0: {6, 7, 0, 1, 2, 3, 4, 5, 9}, 1: set()}, "cuda")
    ids, sc = engine.topk_exact(ue.cuda(), ie.cuda(), uids, 4, excl)
    assert ids.cpu().tolist() == [[8, 0, 1, 2], [6, 7, 0, 1]]
    assert sc[0, 1:].cpu().tolist() == [float("-inf")] * 3
    want_ids, _ = O.recommend_exact(ue, ie, uids, 4, {0: {6, 7, 0, 1, 2, 3, 4, 5, 9}})
    assert torch.equal(ids.cpu(), want_ids)
    # empty batch
    ids, sc = engine.topk_exact(ue.cuda(), ie.cuda(), torch.zeros(0, dtype=torch.int64), 3)
    assert tuple(ids.shape) == (0, 3)


def test_api_errors_and_dims(hnm_lib):
    from hnm_recommendation_b200 import LightGCN
    m = LightGCN(20, 10, embedding_dim=12, num_layers=2, top_k=3).to("cuda")   # dim 12 -> generic kernel
    with pytest.raises(RuntimeError, match="Graph not set"):
        m.recommend(torch.tensor([0]))
    u = torch.randint(0, 20, (60,))
    i = torch.randint(0, 10, (60,)) + 20
    ei = torch.stack([torch.cat([u, i]), torch.cat([i, u])])
    m.set_graph(ei)
    orc = O.LightGCNOracle(20, 10, 12, 2, 3, weight=m.embeddings.weight.detach().cpu())
    orc.set_graph(ei)
    assert_close(m.forward()[0], orc.forward()[0], what="dim12 users")
    assert_close(m.forward()[1], orc.forward()[1], what="dim12 items")
    with pytest.raises(RuntimeError, match="out of range"):
        m.recommend(torch.tensor([0]), k=11)
    with pytest.raises(IndexError):
        m.predict_all_items(torch.tensor([20]))
    with pytest.raises(IndexError):
        m.predict(torch.tensor([0]), torch.tensor([10]))
    assert m.recommend(torch.tensor([3, 4]), k=10).shape == (2, 10)
    assert not m.training                                        # recommend() leaves the module in eval mode
    with pytest.raises(Exception):
        m.set_graph(torch.tensor([[0, 99], [99, 0]]))            # node id out of range


def test_large_k_falls_back_to_device_sort(hnm_lib):
    from hnm_recommendation_b200 import LightGCN
    m = LightGCN(16, 400, embedding_dim=16, top_k=300).to("cuda")
    u = torch.randint(0, 16, (500,)); i = torch.randint(0, 400, (500,)) + 16
    m.set_graph(torch.stack([torch.cat([u, i]), torch.cat([i, u])]))
    got = m.recommend(torch.arange(16))
    ue, ie = m.forward()
    s = O.exact_scores_fp64(ue.cpu(), ie.cpu(), torch.arange(16))
    assert_topk_matches_scores(got, s, 300, rel_tol=1e-5)


def test_user_sharded_propagate_emulated_ranks(hnm_lib):
    """engine.propagate_user_sharded, one emulated rank at a time on one GPU: the all-reduce callback
    checks this rank's partial item sums against the oracle and then substitutes the full sums the
    other ranks would have contributed; the rows the rank owns must come out right."""
    from hnm_recommendation_b200 import engine
    from hnm_recommendation_b200.dist import even_ranges
    U, I, d, L, G = 203, 97, 64, 3, 3
    gen = torch.Generator().manual_seed(3)
    u = torch.randint(0, U, (2500,), generator=gen)
    i = torch.randint(0, I, (2500,), generator=gen) + U
    ei = torch.stack([torch.cat([u, i]), torch.cat([i, u])])
    ew = torch.rand(2500, generator=gen) + 0.5
    ew = torch.cat([ew, ew])
    w = torch.randn(U + I, d, generator=gen) * 0.1
    alphas = O.layer_weights(L)
    for weights in (None, ew):
        rowptr, col, val, dis = O.build_norm_adj(ei, weights, U + I)
        layers = [w]
        for _ in range(L):
            layers.append(O.lightgcn_oracle._spmm(rowptr, col, val, layers[-1]))
        final = sum(a * e for a, e in zip(alphas, layers))
        graph = engine.build_graph(ei, weights, U + I, torch.device("cuda"))
        # raw (un-normalised) adjacency user->item block for the partial sums: W[i, u] = sum of edge weights
        wts = torch.ones(ei.size(1)) if weights is None else weights
        sel = ei[0] >= U                                         # rows that are items
        W = torch.zeros(I, U).index_put_((ei[0][sel] - U, ei[1][sel]), wts[sel], accumulate=True)
        for rank, (u0, u1) in enumerate(even_ranges(U, G)):
            shard = engine.make_user_shard(graph, U, I, u0, u1)
            state = {"layer": 0}

            def allreduce(part):
                xs = dis.unsqueeze(1) * layers[state["layer"]]   # what the kernels gather from
                mine = W[:, u0:u1] @ xs[u0:u1]
                assert_close(part, mine, rtol=1e-5, atol_scale=1e-5, what=f"partial sums rank {rank}")
                part.copy_((W @ xs[:U]).cuda())
                state["layer"] += 1

            acc = engine.propagate_user_sharded(graph, shard, w.cuda(), alphas, L, U, allreduce).cpu()
            assert_close(acc[U:], final[U:], what=f"items, rank {rank}")
            assert_close(acc[u0:u1], final[u0:u1], what=f"own users, rank {rank}")


@pytest.mark.parametrize("dim,weighted", [(32, False), (128, True), (256, False), (64, True), (20, True)])
def test_propagate_dims_and_long_rows(hnm_lib, dim, weighted):
    """Every template instantiation of the propagate kernels (d = 32 / 64 / 128 / 256 and the generic one),
    with rows in all three length classes: warp-per-row, whole-CTA (> 1024 entries) and CTA-cluster (> 8192)."""
    from hnm_recommendation_b200 import LightGCN
    U, I, L = 12000, 40, 2
    gen = torch.Generator().manual_seed(dim)
    # item 0 is bought by every user (12 000 entries), items 1-3 by every 4th user (3 000), the rest at random
    u = torch.cat([torch.arange(U), torch.arange(0, U, 4).repeat(3), torch.randint(0, U, (5000,), generator=gen)])
    i = torch.cat([torch.zeros(U, dtype=torch.long), torch.arange(1, 4).repeat_interleave(U // 4),
                   torch.randint(4, I, (5000,), generator=gen)]) + U
    ei = torch.stack([torch.cat([u, i]), torch.cat([i, u])])
    ew = None
    if weighted:
        half = torch.rand(u.numel(), generator=gen) + 0.5
        ew = torch.cat([half, half])
    torch.manual_seed(1000 + dim)                          # the table's Xavier init: same numbers on every run
    m = LightGCN(U, I, embedding_dim=dim, num_layers=L).to("cuda").eval()
    m.set_graph(ei, ew)
    assert m.graph.num_huge >= 1 and m.graph.num_heavy > m.graph.num_huge
    # the fp64 twin of the oracle is the yardstick here: a row of 12 000 entries summed sequentially in fp32 (the
    # fp32 oracle) is itself ~5e-6 away from the exact sum, so fp32-vs-fp32 at rtol 1e-5 is a coin toss on such rows
    orc = O.LightGCNOracle(U, I, dim, L, weight=m.embeddings.weight.detach().cpu(), dtype=torch.float64)
    orc.set_graph(ei, ew)
    gu, gi = m.forward()
    ou, oi = orc.forward()
    assert_close(gu, ou, what=f"users d={dim}")
    assert_close(gi, oi, what=f"items d={dim}")
    # deterministic: the long rows are summed in a fixed order
    m.cache_embeddings = False
    gu2, gi2 = m.forward()
    assert torch.equal(gu, gu2) and torch.equal(gi, gi2)


@pytest.mark.parametrize("dim,weighted,chunks", [(64, False, 3), (64, True, 5), (128, False, 2), (32, True, 4)])
def test_propagate_item_rows_in_user_chunks(hnm_lib, dim, weighted, chunks):
    """The single-GPU form that walks the item rows once per chunk of users (hnm_lightgcn_partial with
    accumulate + hnm_lightgcn_finish): same embeddings as the oracle and as the one-pass kernels, for rows of
    every length class, and deterministic."""
    from hnm_recommendation_b200 import LightGCN, engine
    U, I, L = 30000, 40, 3
    gen = torch.Generator().manual_seed(dim + chunks)
    u = torch.cat([torch.arange(U), torch.arange(0, U, 4).repeat(3), torch.randint(0, U, (5000,), generator=gen)])
    i = torch.cat([torch.zeros(U, dtype=torch.long), torch.arange(1, 4).repeat_interleave(U // 4),
                   torch.randint(4, I, (5000,), generator=gen)]) + U
    ei = torch.stack([torch.cat([u, i]), torch.cat([i, u])])
    ew = None
    if weighted:
        half = torch.rand(u.numel(), generator=gen) + 0.5
        ew = torch.cat([half, half])
    m = LightGCN(U, I, embedding_dim=dim, num_layers=L).to("cuda")
    m.set_graph(ei, ew)
    assert m._item_chunks is None                                  # the user block fits the L2: one pass
    assert engine.is_bipartite(m.graph, U)
    one_pass = engine.propagate(m.graph, m.embeddings.weight, m.alpha, L)
    ic = engine.make_item_chunks(m.graph, U, I, dim, num_chunks=chunks)
    assert ic is not None and len(ic.chunks) == chunks
    # item 0 (bought by every user) is a cluster-kernel row in a chunk of >= 8192 users, a whole-CTA row otherwise
    assert all(sh.heavy_rows.numel() >= 1 for sh in ic.chunks)
    assert any(sh.num_huge for sh in ic.chunks) == (U // chunks > 8192)
    chunked = engine.propagate(m.graph, m.embeddings.weight, m.alpha, L, item_chunks=ic)
    orc = O.LightGCNOracle(U, I, dim, L, weight=m.embeddings.weight.detach().cpu())
    orc.set_graph(ei, ew)
    ou, oi = orc.forward()
    assert_close(chunked[:U], ou, what="users, chunked")
    # item 0's row is a 30 000-term fp32 sum with cancellation: entries near zero carry ~sqrt(n) ulp of the terms
    assert_close(chunked[U:], oi, atol_scale=1e-5, what="items, chunked")
    assert_close(chunked[:U], one_pass[:U], what="users, chunked vs one pass")
    assert_close(chunked[U:], one_pass[U:], atol_scale=1e-5, what="items, chunked vs one pass")
    again = engine.propagate(m.graph, m.embeddings.weight, m.alpha, L, item_chunks=ic)
    assert torch.equal(chunked, again)                             # chunk order is fixed: deterministic


def test_non_bipartite_graph_is_detected_and_not_chunked(hnm_lib):
    """A user-user edge: still a valid graph for set_graph/forward, but neither chunked nor user-partitioned."""
    from hnm_recommendation_b200 import LightGCN, engine
    U, I = 50, 30
    gen = torch.Generator().manual_seed(1)
    u = torch.randint(0, U, (400,), generator=gen)
    i = torch.randint(0, I, (400,), generator=gen) + U
    ei = torch.stack([torch.cat([u, i, torch.tensor([3])]), torch.cat([i, u, torch.tensor([7])])])
    m = LightGCN(U, I, embedding_dim=16, num_layers=2).to("cuda")
    m.set_graph(ei)
    assert not engine.is_bipartite(m.graph, U)
    assert engine.make_item_chunks(m.graph, U, I, 16, num_chunks=4) is None
    orc = O.LightGCNOracle(U, I, 16, 2, weight=m.embeddings.weight.detach().cpu())
    orc.set_graph(ei)
    gu, gi = m.forward()
    ou, oi = orc.forward()
    assert_close(gu, ou, what="users")
    assert_close(gi, oi, what="items")


@pytest.mark.parametrize("symmetric,weighted,alpha", [(True, False, None), (True, True, 0.5), (False, True, None)])
def test_bpr_loss_gradients_match_autograd_oracle(hnm_lib, symmetric, weighted, alpha):
    """forward_with_grad(): the backward pass is the propagate kernels on the transposed adjacency
    (dL/dE0 = sum_l alpha_l (A_hat^T)^l dL/dfinal).  Loss and gradient of bpr_loss (lightgcn.py:206-245) against
    plain autograd through a dense fp64 restatement -- also for an edge list that is NOT symmetric."""
    from hnm_recommendation_b200 import LightGCN
    U, I, d, L, E, B = 180, 70, 64, 3, 1500, 256
    gen = torch.Generator().manual_seed(11)
    u = torch.randint(0, U, (E,), generator=gen)
    i = torch.randint(0, I, (E,), generator=gen) + U
    if symmetric:
        ei = torch.stack([torch.cat([u, i]), torch.cat([i, u])])
        half = torch.rand(E, generator=gen) + 0.5
        ew = torch.cat([half, half]) if weighted else None
    else:
        keep = torch.rand(E, generator=gen) < 0.6                      # only some reverse edges exist
        ei = torch.stack([torch.cat([u, i[keep]]), torch.cat([i, u[keep]])])
        ew = torch.rand(ei.size(1), generator=gen) + 0.5
    m = LightGCN(U, I, embedding_dim=d, num_layers=L, alpha=alpha, weight_decay=1e-2).to("cuda")
    with torch.no_grad():
        m.embeddings.weight.copy_(torch.randn(U + I, d, generator=gen) * 0.3)
    m.set_graph(ei, ew)
    uid = torch.randint(0, U, (B,), generator=gen)
    pos = torch.randint(0, I, (B,), generator=gen)
    neg = torch.randint(0, I, (B,), generator=gen)
    loss = m.training_step({"user_ids": uid.cuda(), "pos_items": pos.cuda(), "neg_items": neg.cuda()}, 0)
    loss.backward()
    grad = m.embeddings.weight.grad
    assert (m._adjoint_graph() is m.graph) == symmetric

    rowptr, col, val, _ = O.build_norm_adj(ei, ew, U + I, dtype=torch.float64)
    w64 = m.embeddings.weight.detach().cpu().double().requires_grad_()
    want = O.bpr_loss(w64, rowptr, col, val, U, L, m.alpha, uid, pos, neg, m.weight_decay)
    want.backward()
    assert abs(float(loss) - float(want)) <= 1e-5 * abs(float(want))
    assert_close(grad, w64.grad, rtol=1e-4, atol_scale=1e-5, what="dL/d embeddings.weight")
    # train mode: forward() is the reference's differentiable forward; eval / no_grad: the cached inference path
    assert m.forward()[0].requires_grad
    with torch.no_grad():
        assert not m.forward()[0].requires_grad
    m.eval()
    assert not m.forward()[0].requires_grad
    # an optimizer step through the reference's configuration changes the weights and invalidates the cache
    before = m.forward()[0].clone()
    opt = m.configure_optimizers()["optimizer"]
    opt.step()
    assert not torch.equal(m.forward()[0], before)


@pytest.mark.parametrize("path", golden_files("train"), ids=lambda p: p.split("train_")[-1][:-4])
def test_training_step_matches_reference_golden(hnm_lib, path):
    """training_step -> bpr_loss -> backward through the propagate kernels, against the loss and gradient of
    the reference's own training_step on the same inputs (tests/golden/make_golden_train.py)."""
    from hnm_recommendation_b200 import LightGCN
    g = load_golden(path)
    alpha = None if np.isnan(g["alpha"]) else float(g["alpha"])
    m = LightGCN(int(g["num_users"]), int(g["num_items"]), embedding_dim=int(g["embedding_dim"]),
                 num_layers=int(g["num_layers"]), alpha=alpha, weight_decay=float(g["weight_decay"]))
    m.load_state_dict({"embeddings.weight": torch.from_numpy(g["weight"])})
    m = m.to("cuda")
    ew = torch.from_numpy(g["edge_weight"]) if g["edge_weight"].size else None
    m.set_graph(torch.from_numpy(g["edge_index"]), ew)
    batch = {k: torch.from_numpy(g[k]).cuda() for k in ("user_ids", "pos_items", "neg_items")}
    loss = m.training_step(batch, 0)
    loss.backward()
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-5)
    assert_close(m.embeddings.weight.grad, g["grad"], rtol=1e-4, atol_scale=1e-5, what="gradient")
