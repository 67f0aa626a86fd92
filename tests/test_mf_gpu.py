"""MatrixFactorization (src/models/matrix_factorization.py): U V^T + b_u + b_i + b_g ranked
through the fused tensor-core kernel (bias as an extra K chunk), against the reference's own outputs
(tests/golden/mf_*.npz) and the oracle."""
import numpy as np
import pytest
import torch

import oracle as O
from conftest import assert_close, filter_dict, golden_files, load_golden

pytestmark = pytest.mark.gpu


def _model(g):
    from hnm_recommendation_b200 import MatrixFactorization
    m = MatrixFactorization(int(g["num_users"]), int(g["num_items"]), embedding_dim=int(g["embedding_dim"]),
                            top_k=int(g["top_k"]), sparse=False)
    state = {k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("state.")}
    assert set(state) == set(m.state_dict())                    # the reference's state_dict keys
    m.load_state_dict(state)
    return m.to("cuda").eval(), state


@pytest.mark.parametrize("path", golden_files("mf"), ids=lambda p: p.split("mf_")[-1][:-4])
def test_mf_golden(hnm_lib, path):
    g = load_golden(path)
    m, state = _model(g)
    uids, iids = torch.from_numpy(g["user_ids"]), torch.from_numpy(g["item_ids"])
    with torch.no_grad():
        assert_close(m(uids, iids), g["pred"], what="forward")
    au = torch.from_numpy(g["all_user_ids"])
    scores = m.predict_all_items(au)
    assert_close(scores, g["scores"], what="predict_all_items")
    # recommend: the reference's fp32 ranking up to near-ties; exactly the oracle's canonical ranking
    k = int(g["top_k"])
    want = O.mf_recommend_exact(state, au, k)
    got = m.recommend(au)
    assert torch.equal(got.cpu(), want)
    g["user_ids"] = g["all_user_ids"]
    filt = filter_dict(g)
    assert torch.equal(m.recommend(au, filter_items=filt).cpu(), O.mf_recommend_exact(state, au, k, filt))
    differ = (got.cpu() != torch.from_numpy(g["topk_canonical"])).any(dim=1).sum()
    assert int(differ) <= 2                                     # fp32 near-ties of the reference's matmul only


def test_mf_fused_path_all_users(hnm_lib):
    """More users than the small-batch threshold: the fused kernel at d' = 128 (64 + the bias chunk)."""
    from hnm_recommendation_b200 import MatrixFactorization
    torch.manual_seed(3)
    U, I = 6000, 4100
    m = MatrixFactorization(U, I, sparse=False).to("cuda").eval()
    with torch.no_grad():
        m.item_bias.weight.normal_(0, 0.01)
        m.user_bias.weight.normal_(0, 0.01)
        m.global_bias.fill_(0.002)
    state = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ids = m.recommend_all()
    assert m._scorer is not None and m._scorer.dim == 128
    uids = torch.arange(0, U, 11)
    assert torch.equal(ids[uids.cuda()].cpu(), O.mf_recommend_exact(state, uids, 12))
    stats = m._scorer.last_stats
    assert stats["tier3"] <= U // 100, stats
    # a parameter update is seen (version counters), a .data write after invalidate()
    with torch.no_grad():
        m.item_bias.weight.add_(torch.randn_like(m.item_bias.weight) * 0.05)
    state = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    assert torch.equal(m.recommend(uids).cpu(), O.mf_recommend_exact(state, uids, 12))
    # training step (BCE with logits, matrix_factorization.py:133-156)
    m.train()
    loss = m.training_step({"user_ids": torch.randint(0, U, (64,)), "item_ids": torch.randint(0, I, (64,)),
                            "labels": torch.randint(0, 2, (64,))}, 0)
    loss.backward()
    assert m.item_bias.weight.grad is not None and m.global_bias.grad is not None
    assert isinstance(m.configure_optimizers(), torch.optim.Adam)
