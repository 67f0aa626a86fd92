"""Size-independent properties at BASELINE.json's full H&M shape (1 371 980 x 105 542 x 31 788 324, d 64):
the oracle cannot run here in seconds, so the CUDA path is checked through invariants of the maths."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hm_model(hnm_lib):
    from hnm_recommendation_b200 import LightGCN, synth
    data = synth.interactions(synth.HM_USERS, synth.HM_ITEMS, synth.HM_EDGES, seed=42)
    m = LightGCN(data.num_users, data.num_items).to("cuda")
    with torch.no_grad():
        m.embeddings.weight.copy_(synth.xavier_embeddings(data.num_users + data.num_items, 64, seed=42))
    m.set_graph(data.edge_index().cuda())
    m.cache_embeddings = False
    return m


def test_graph_invariants(hm_model):
    g = hm_model.graph
    n = hm_model.num_users + hm_model.num_items
    assert g.nnz == 2 * 31_788_324 + n
    rp = g.rowptr.long()
    assert int(rp[0]) == 0 and int(rp[-1]) == g.nnz and bool((rp[1:] > rp[:-1]).all())      # self loop in every row
    deg = (rp[1:] - rp[:-1]).float()
    assert torch.allclose(g.dis, deg.rsqrt(), rtol=2e-7, atol=0)
    # columns sorted inside every row (checked on a sample of rows, including the longest one)
    rows = torch.cat([torch.randint(0, n, (2000,), device="cuda"), deg.argmax().view(1)])
    for r in rows.tolist()[-50:]:
        c = g.col[int(rp[r]):int(rp[r + 1])]
        assert bool((c[1:] >= c[:-1]).all())
    # symmetric: entry count of user->item equals item->user
    assert int(rp[hm_model.num_users]) - hm_model.num_users == g.nnz - int(rp[hm_model.num_users]) - hm_model.num_items


def test_propagate_fixed_point_and_linearity(hm_model):
    """A_hat = D^-1/2 (A+I) D^-1/2 has eigenvector sqrt(deg) with eigenvalue 1, so an embedding table whose
    columns are all sqrt(deg) must come back unchanged from every layer and from the layer mean."""
    m = hm_model
    g = m.graph
    rp = g.rowptr.long()
    sq = (rp[1:] - rp[:-1]).float().sqrt()
    w_saved = m.embeddings.weight.detach().clone()
    try:
        with torch.no_grad():
            m.embeddings.weight.copy_(sq.unsqueeze(1).expand(-1, 64))
        ue, ie = m.forward()
        got = torch.cat([ue, ie])
        rel = ((got - sq.unsqueeze(1)).abs() / sq.unsqueeze(1)).max()
        assert float(rel) < 1e-5, float(rel)
        # linearity: f(a x + b y) = a f(x) + b f(y)
        x = torch.randn_like(w_saved) * 0.05
        with torch.no_grad():
            m.embeddings.weight.copy_(w_saved)
        fx = torch.cat(m.forward()).clone()
        with torch.no_grad():
            m.embeddings.weight.copy_(x)
        fy = torch.cat(m.forward()).clone()
        with torch.no_grad():
            m.embeddings.weight.copy_(2.0 * w_saved - 0.5 * x)
        fz = torch.cat(m.forward())
        want = 2.0 * fx - 0.5 * fy
        scale = float(want.abs().max())
        assert float((fz - want).abs().max()) < 2e-5 * scale
    finally:
        with torch.no_grad():
            m.embeddings.weight.copy_(w_saved)


def test_recommend_all_properties_and_spot_check(hm_model):
    from hnm_recommendation_b200 import engine
    m = hm_model
    ids, sc = m.recommend_all(return_scores=True)
    U, I = m.num_users, m.num_items
    assert tuple(ids.shape) == (U, 12) and ids.dtype == torch.int64
    assert int(ids.min()) >= 0 and int(ids.max()) < I
    assert bool((sc[:, 1:] <= sc[:, :-1]).all())                                  # score descending
    tie = sc[:, 1:] == sc[:, :-1]
    assert bool((ids[:, 1:][tie] > ids[:, :-1][tie]).all())                       # ties by id ascending
    srt = ids.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                                # no duplicates in a row
    stats = m._scorer.last_stats
    assert stats["tier3"] <= U // 1000, stats                                      # exact fallback stays rare
    # bit-exact agreement with the brute-force fp64 kernel on a random sample of users
    ue, ie = m.forward()
    uids = torch.randint(0, U, (4096,), device="cuda")
    e_ids, e_sc = engine.topk_exact(ue, ie, uids, 12)
    assert torch.equal(ids[uids], e_ids) and torch.equal(sc[uids], e_sc)
    # the 12th score dominates the scores of random other items (fp64 recomputation)
    probe = torch.randint(0, I, (4096, 64), device="cuda")
    s = torch.einsum("bd,bkd->bk", ue[uids].double(), ie[probe].double())
    chosen = (probe.unsqueeze(2) == ids[uids].unsqueeze(1)).any(dim=2)
    assert bool((s[~chosen] <= sc[uids][:, -1:].expand(-1, 64)[~chosen] + 1e-12).all())
    # single-user API path agrees with the all-users path
    some = uids[:257]
    assert torch.equal(m.recommend(some), ids[some])
