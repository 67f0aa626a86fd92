"""The plain-C restatement (oracle/c/hnm_oracle.c) against the golden vectors produced by the reference's own
files and against the PyTorch oracle: two independent CPU statements of the path must agree with each other and
with the reference before either is trusted as the checker of the CUDA path."""
import numpy as np
import pytest
import torch

import oracle as O
from oracle import c_oracle as CO
from conftest import assert_close, filter_dict, golden_files, load_golden

LG = golden_files("lightgcn")
NCF = golden_files("ncf")


def _graph_inputs(g):
    ew = g["edge_weight"] if g["edge_weight"].size else None
    return g["edge_index"], ew, int(g["num_users"]) + int(g["num_items"])


@pytest.mark.parametrize("path", LG, ids=lambda p: p.split("lightgcn_")[-1][:-4])
def test_c_graph_and_forward_match_reference_golden(path):
    g = load_golden(path)
    ei, ew, n = _graph_inputs(g)
    rowptr, col, val, dis = CO.norm_adj(ei, ew, n)
    t_rowptr, t_col, t_val, t_dis = O.build_norm_adj(torch.from_numpy(ei), None if ew is None else torch.from_numpy(ew), n)
    assert np.array_equal(rowptr, t_rowptr.numpy()) and np.array_equal(col, t_col.numpy())      # index work: exact
    assert_close(dis, t_dis, rtol=2e-7, atol_scale=0.0, what="dis")                             # 1/sqrt vs pow(-0.5): 1 ulp
    assert_close(val, t_val, rtol=4e-7, atol_scale=0.0, what="val")
    ue, ie = CO.forward(g["weight"], rowptr, col, val, int(g["num_users"]), int(g["num_layers"]), g["alpha_list"])
    assert_close(ue, g["user_emb"], what="user_emb vs the reference")                           # P1: rtol 1e-5
    assert_close(ie, g["item_emb"], what="item_emb vs the reference")


@pytest.mark.parametrize("path", LG, ids=lambda p: p.split("lightgcn_")[-1][:-4])
def test_c_topk_bit_identical_to_torch_oracle_and_reference_lists(path):
    g = load_golden(path)
    ue, ie = torch.from_numpy(g["user_emb"]), torch.from_numpy(g["item_emb"])
    uids = torch.from_numpy(g["user_ids"])
    k = int(g["top_k"])
    for filt in (None, filter_dict(g)):
        ids, sc = CO.topk_exact(ue, ie, uids, k, filt)
        w_ids, w_sc = O.recommend_exact(ue, ie, uids, k, filt)
        assert np.array_equal(ids, w_ids.numpy()) and np.array_equal(sc, w_sc.numpy())          # ids AND fp64 bits
    # the reference's own lists (fp32 sgemm + torch.topk) agree except at near-ties
    ids, _ = CO.topk_exact(ue, ie, uids, k)
    assert int((ids != g["topk_canonical"]).any(axis=1).sum()) <= max(1, len(uids) // 20)


def test_c_topk_ties_filter_tail_and_errors():
    ue = np.array([[1.0, 0.0, 2.0, 0.5]] * 2, np.float32)
    ie = np.array([[1.0, 1.0, 1.0, 1.0]] * 6 + [[2.0, 2.0, 2.0, 2.0]] * 2 + [[-1.0, 0, 0, 0]] * 2, np.float32)
    ids, _ = CO.topk_exact(ue, ie, [0, 1], 5)
    assert ids.tolist() == [[6, 7, 0, 1, 2]] * 2                                                # exact ties: id ascending
    ids, sc = CO.topk_exact(ue, ie, [0], 10, {0: {6, 7, 0, 1, 2, 3, 4}})
    assert ids[0, :3].tolist() == [5, 8, 9] and ids[0, 3:].tolist() == [0, 1, 2, 3, 4, 6, 7]    # -inf tail, ids ascending
    assert np.isinf(sc[0, 3:]).all()
    w_ids, w_sc = O.recommend_exact(torch.from_numpy(ue), torch.from_numpy(ie), torch.tensor([0]), 10,
                                    {0: {6, 7, 0, 1, 2, 3, 4}})
    assert np.array_equal(ids, w_ids.numpy()) and np.array_equal(sc, w_sc.numpy())
    with pytest.raises(RuntimeError, match="out of range"):
        CO.topk_exact(ue, ie, [0], 11)


def test_c_isolated_node_and_duplicates():
    # node 3 has no edge: its self loop gives deg 1 (src/models/lightgcn.py:127-132); (0, 2) appears twice
    ei = np.array([[0, 2, 0, 2, 1, 2], [2, 0, 2, 0, 2, 1]], np.int64)
    rowptr, col, val, dis = CO.norm_adj(ei, None, 4)
    assert rowptr.tolist() == [0, 3, 5, 9, 10] and col.tolist() == [0, 2, 2, 1, 2, 0, 0, 1, 2, 3]
    assert dis[3] == 1.0 and val[-1] == 1.0
    assert np.isclose(dis[0], 3 ** -0.5) and np.isclose(dis[2], 0.5)


@pytest.mark.parametrize("path", NCF, ids=lambda p: p.split("ncf_")[-1][:-4])
def test_c_ncf_forward_matches_reference_golden(path):
    g = load_golden(path)
    state = {k[len("state."):]: v for k, v in g.items() if k.startswith("state.")}
    out = CO.ncf_forward(state, g["user_ids"], g["item_ids"])
    assert_close(out, g["logits"], what="logits vs the reference")                              # P4: rtol 1e-5
    t_state = {k: torch.from_numpy(v) for k, v in state.items()}
    t_out = O.ncf_forward(t_state, torch.from_numpy(g["user_ids"]), torch.from_numpy(g["item_ids"]))
    assert_close(out, t_out, what="logits vs the PyTorch oracle")
    nu = len(g["all_user_ids"])
    ni = int(g["num_items"])
    uu = np.repeat(g["all_user_ids"], ni)
    ii = np.tile(np.arange(ni), nu)
    assert_close(CO.ncf_forward(state, uu, ii).reshape(nu, ni), g["all_scores"], what="all_scores vs the reference")


@pytest.mark.parametrize("path", golden_files("mf"), ids=lambda p: p.split("mf_")[-1][:-4])
def test_c_mf_ranking_matches_torch_oracle_and_reference_lists(path):
    """MatrixFactorization (src/models/matrix_factorization.py:108-131,217-245): exact dot product + b_i decides."""
    g = load_golden(path)
    state = {k[len("state."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("state.")}
    uids = torch.from_numpy(g["all_user_ids"])
    k = int(g["top_k"])
    filt = {}
    for r, i in zip(g["filter_rows"].tolist(), g["filter_items"].tolist()):
        filt.setdefault(int(g["all_user_ids"][r]), set()).add(int(i))
    for f in (None, filt):
        ids, sc = CO.topk_exact(state["user_embeddings.weight"], state["item_embeddings.weight"], uids, k, f,
                                item_bias=state["item_bias.weight"])
        want = O.mf_recommend_exact(state, uids, k, f)
        assert np.array_equal(ids, want.numpy())
        s64 = O.mf_rank_scores_fp64(state, uids)
        assert np.array_equal(sc, torch.gather(s64, 1, want).numpy())                           # the fp64 bits too
    ids, _ = CO.topk_exact(state["user_embeddings.weight"], state["item_embeddings.weight"], uids, k,
                           item_bias=state["item_bias.weight"])
    assert int((ids != g["topk_canonical"]).any(axis=1).sum()) <= max(1, len(uids) // 10)       # near-ties only
