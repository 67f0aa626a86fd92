"""The oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import oracle as O
from conftest import assert_close, filter_dict, golden_files, load_golden

LG = golden_files("lightgcn")
NCF = golden_files("ncf")


def _oracle_from(g, dtype=torch.float32):
    alpha = None if np.isnan(g["alpha"]) else float(g["alpha"])
    m = O.LightGCNOracle(int(g["num_users"]), int(g["num_items"]), int(g["embedding_dim"]),
                         int(g["num_layers"]), int(g["top_k"]), alpha,
                         weight=torch.from_numpy(g["weight"]), dtype=dtype)
    ew = torch.from_numpy(g["edge_weight"]) if g["edge_weight"].size else None
    m.set_graph(torch.from_numpy(g["edge_index"]), ew)
    return m


def test_golden_present():
    assert len(LG) >= 7 and len(NCF) >= 3


@pytest.mark.parametrize("path", LG, ids=lambda p: p.split("lightgcn_")[-1][:-4])
def test_lightgcn_oracle_matches_reference(path):
    g = load_golden(path)
    m = _oracle_from(g)
    assert m.alpha == pytest.approx(g["alpha_list"].tolist(), rel=0, abs=0)
    ue, ie = m.forward()
    assert_close(ue, g["user_emb"], what="user_emb")
    assert_close(ie, g["item_emb"], what="item_emb")
    uids = torch.from_numpy(g["user_ids"])
    assert_close(m.predict_all_items(uids), g["scores"], what="scores")
    assert_close(m.predict(uids, torch.from_numpy(g["item_ids"])), g["pair_scores"], what="pair")


@pytest.mark.parametrize("path", LG, ids=lambda p: p.split("lightgcn_")[-1][:-4])
def test_lightgcn_topk_matches_reference(path):
    g = load_golden(path)
    k = int(g["top_k"])
    ue, ie = torch.from_numpy(g["user_emb"]), torch.from_numpy(g["item_emb"])
    uids = torch.from_numpy(g["user_ids"])
    # fed with the reference's own embeddings: canonical top-k must be identical
    got = O.recommend(ue, ie, uids, k)
    assert torch.equal(got, torch.from_numpy(g["topk_canonical"]))
    got_f = O.recommend(ue, ie, uids, k, filter_dict(g))
    assert torch.equal(got_f, torch.from_numpy(g["topk_filtered_canonical"]))
    # torch.topk (what the reference returned) agrees wherever scores are tie-free
    s = torch.from_numpy(g["scores"])
    top = torch.gather(s, 1, torch.from_numpy(g["topk_canonical"]))
    tie_free = (top[:, 1:] != top[:, :-1]).all(dim=1)
    assert torch.equal(torch.from_numpy(g["recommend_raw"])[tie_free], got[tie_free])
    # fp64 sequential-k scores order the same items unless two fp32 scores are within rounding
    ex_ids, ex_s = O.recommend_exact(ue, ie, uids, k)
    diff = ex_ids != got
    if diff.any():
        rel_gap = (ex_s[:, :-1] - ex_s[:, 1:]).abs() / ex_s.abs().max()
        assert float(rel_gap[diff[:, :-1] | diff[:, 1:]].min()) < 1e-6


def test_fp64_twin_close_to_fp32():
    g = load_golden([p for p in LG if p.endswith("mid.npz")][0])
    u32, i32 = _oracle_from(g).forward()
    u64, i64 = _oracle_from(g, torch.float64).forward()
    assert_close(u32, u64, what="fp32 vs fp64 users")
    assert_close(i32, i64, what="fp32 vs fp64 items")


def test_topk_tie_break_is_id_ascending():
    s = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0, 0.0]])
    assert O.topk_canonical(s, 4).tolist() == [[1, 2, 4, 3]]
    with pytest.raises(RuntimeError):
        O.topk_canonical(s, 7)


def test_forward_before_set_graph_raises():
    m = O.LightGCNOracle(4, 3, 8, 2)
    with pytest.raises(RuntimeError, match="Graph not set"):
        m.forward()


def test_isolated_node_keeps_self_loop():
    # node with no edges: deg = 1 from the self loop, so A_hat row = e_i and every layer returns E0_i
    m = O.LightGCNOracle(3, 2, 4, 3, weight=torch.arange(20.0).view(5, 4))
    m.set_graph(torch.tensor([[0, 3], [3, 0]]))
    u, i = m.forward()
    assert_close(u[1], m.weight[1], what="isolated user")
    assert_close(i[1], m.weight[4], what="isolated item")


@pytest.mark.parametrize("path", NCF, ids=lambda p: p.split("ncf_")[-1][:-4])
def test_ncf_oracle_matches_reference(path):
    g = load_golden(path)
    state = {k[len("state."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("state.")}
    m = O.NeuralCFOracle(int(g["num_users"]), int(g["num_items"]), state=state)
    got = m.forward(torch.from_numpy(g["user_ids"]), torch.from_numpy(g["item_ids"]))
    assert_close(got, g["logits"], rtol=1e-5, atol_scale=1e-6, what="logits")
    one = m.forward(torch.from_numpy(g["user_ids"][:1]), torch.from_numpy(g["item_ids"][:1]))
    assert one.dim() == 0 and g["single_logit"].ndim == 0          # .squeeze() quirk, neural_cf.py:139
    alls = m.predict_all_items(torch.from_numpy(g["all_user_ids"]))
    assert_close(alls, g["all_scores"], what="all_scores")
    assert torch.equal(O.topk_canonical(torch.from_numpy(g["all_scores"]), int(g["top_k"])),
                       torch.from_numpy(g["topk_canonical"]))


TRAIN = golden_files("train")


@pytest.mark.parametrize("path", TRAIN, ids=lambda p: p.split("train_")[-1][:-4])
def test_bpr_loss_oracle_matches_reference(path):
    """oracle.bpr_loss (dense fp64 restatement + plain autograd) against the loss and the gradient the
    reference's own LightGCN.training_step produced (tests/golden/make_golden_train.py)."""
    g = load_golden(path)
    U, L = int(g["num_users"]), int(g["num_layers"])
    n = U + int(g["num_items"])
    alpha = None if np.isnan(g["alpha"]) else float(g["alpha"])
    ew = torch.from_numpy(g["edge_weight"]) if g["edge_weight"].size else None
    rowptr, col, val, _ = O.build_norm_adj(torch.from_numpy(g["edge_index"]), ew, n, dtype=torch.float64)
    w = torch.from_numpy(g["weight"]).double().requires_grad_()
    loss = O.bpr_loss(w, rowptr, col, val, U, L, O.layer_weights(L, alpha), torch.from_numpy(g["user_ids"]),
                      torch.from_numpy(g["pos_items"]), torch.from_numpy(g["neg_items"]), float(g["weight_decay"]))
    loss.backward()
    assert len(TRAIN) >= 2
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-5)
    assert_close(w.grad, g["grad"], rtol=1e-4, atol_scale=1e-5, what="gradient")


@pytest.mark.parametrize("path", golden_files("mf"), ids=lambda p: p.split("mf_")[-1][:-4])
def test_mf_oracle_matches_reference_golden(path):
    """oracle/mf_oracle.py against the outputs of the reference's own matrix_factorization.py."""
    g = load_golden(path)
    state = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("state.")}
    uids, iids = torch.from_numpy(g["user_ids"]), torch.from_numpy(g["item_ids"])
    assert torch.allclose(O.mf_forward(state, uids, iids), torch.from_numpy(g["pred"]), rtol=1e-6, atol=1e-9)
    au = torch.from_numpy(g["all_user_ids"])
    scores = O.mf_predict_all_items(state, au)
    assert torch.allclose(scores, torch.from_numpy(g["scores"]), rtol=1e-6, atol=1e-9)
    k = int(g["top_k"])
    canon = O.mf_recommend_exact(state, au, k)
    # the exact ranking differs from the fp32 one only at near-ties
    assert int((canon != torch.from_numpy(g["topk_canonical"])).any(dim=1).sum()) <= 2
