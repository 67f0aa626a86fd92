"""GPU parity: tensor-core score+select path (bit-exact top-k through certificate/fallback) and NeuralCF."""
import numpy as np
import pytest
import torch

import oracle as O
from conftest import assert_close, golden_files, load_golden

pytestmark = pytest.mark.gpu


def _emb(u, i, seed, kind="randn"):
    g = torch.Generator().manual_seed(seed)
    if kind == "randn":
        return torch.randn(u, 64, generator=g) * 0.1, torch.randn(i, 64, generator=g) * 0.1
    # tiny magnitudes like a Xavier-initialised 1.5M-row table, with a common mean (propagated embeddings)
    base = torch.randn(1, 64, generator=g) * 2e-3
    return (torch.randn(u, 64, generator=g) * 1e-3 + base), (torch.randn(i, 64, generator=g) * 1e-3 + base)


@pytest.mark.parametrize("u,i,kind", [(700, 1000, "randn"), (1537, 5000, "small"), (512, 256, "randn"),
                                      (130, 4100, "small")])
def test_fused_topk_bit_exact(hnm_lib, u, i, kind):
    from hnm_recommendation_b200.scorer import FusedScorer
    ue, ie = _emb(u, i, seed=u + i, kind=kind)
    want_ids, want_s = O.recommend_exact(ue, ie, torch.arange(u), 12)
    sc = FusedScorer(ue.cuda(), ie.cuda())
    ids, s = sc.topk(None, 12)
    assert torch.equal(ids.cpu(), want_ids)
    assert torch.equal(s.cpu(), want_s)
    # the tensor-core path alone must already certify nearly everyone on generic data
    assert sc.last_stats["uncertified"] <= max(2, u // 100), sc.last_stats
    # without the fallback, certified rows are exact and uncertified rows are flagged, never silently wrong
    ids2, _ = sc.topk(None, 12, fallback=False)
    same = (ids2.cpu() == want_ids).all(dim=1)
    assert int((~same).sum()) <= sc.last_stats["uncertified"] or sc.last_stats["uncertified"] == 0


def test_fused_subset_filter_and_shard(hnm_lib):
    from hnm_recommendation_b200.scorer import FusedScorer
    ue, ie = _emb(900, 3000, seed=5)
    uids = torch.randperm(900)[:333]
    filt = {int(x): set(torch.randint(0, 3000, (40,)).tolist()) for x in uids[::2].tolist()}
    # make sure some filtered items are the user's true best ones
    top = O.recommend_exact(ue, ie, uids, 12)[0]
    for r, x in enumerate(uids.tolist()):
        if x in filt:
            filt[x].update(top[r, :5].tolist())
    want_ids, want_s = O.recommend_exact(ue, ie, uids, 12, filt)
    sc = FusedScorer(ue.cuda(), ie.cuda())
    ids, s = sc.topk(uids, 12, filt)
    assert torch.equal(ids.cpu(), want_ids) and torch.equal(s.cpu(), want_s)
    # item shard [1000, 3000): ids are global
    shard = FusedScorer(ue.cuda(), ie[1000:].cuda().contiguous(), item_begin=1000)
    ids, s = shard.topk(uids, 7)
    w_ids, w_s = O.recommend_exact(ue, ie[1000:], uids, 7)
    assert torch.equal(ids.cpu(), w_ids + 1000) and torch.equal(s.cpu(), w_s)


@pytest.mark.parametrize("u,i,kind", [(300, 16500, "randn"), (1000, 33000, "small"), (-9, 8300, "randn")])
def test_fused_sliced_left_over_tiles_bit_exact(hnm_lib, u, i, kind):
    """User tiles that do not fill a whole pass of the persistent grid have their item range sliced over
    the CTAs (one candidate list per slice, merged under a recomputed threshold): same lists, bit for bit.
    u < 0: one whole pass on every CTA plus -u left-over tiles (the count depends on the CTA shape in use)."""
    import ctypes as C
    from hnm_recommendation_b200 import engine
    from hnm_recommendation_b200.scorer import FusedScorer
    plan = (C.c_int32 * 6)()
    pad = lambda x: (x + 127) // 128 * 128
    if u < 0:
        assert hnm_lib.hnm_score_topk_fused_plan(128, pad(i), plan) == 0
        u = (plan[0] * plan[5] - u) * 128 - 50
    assert hnm_lib.hnm_score_topk_fused_plan(pad(u), pad(i), plan) == 0
    grid, full, tile0, triples, slices, mu = list(plan)
    assert triples > 0 and slices > 1, list(plan)                  # the case under test
    ue, ie = _emb(u, i, seed=u + i, kind=kind)
    ue, ie = ue.cuda(), ie.cuda()
    sc = FusedScorer(ue, ie)
    ids, s = sc.topk(None, 12)
    assert sc.last_stats["uncertified"] <= max(3, u // 50), sc.last_stats
    # the sliced users are the last ones; check all of them (and a sample of the rest) against brute force
    first_sliced = min(u, tile0 * 128)
    check = torch.cat([torch.arange(first_sliced, u), torch.randperm(max(first_sliced, 1))[:256] % u]).cuda()
    w_ids, w_s = engine.topk_exact(ue, ie, check, 12)
    assert torch.equal(ids[check], w_ids) and torch.equal(s[check], w_s)
    # a filter that removes every sliced user's best items
    some = check[:200].tolist()
    filt = {x: set(w_ids[r, :6].tolist()) for r, x in enumerate(some)}
    f_ids, f_s = sc.topk(check[:200], 12, filt)
    e_ids, e_s = engine.topk_exact(ue, ie, check[:200], 12, engine.exclusion_csr(check[:200], filt, ue.device))
    assert torch.equal(f_ids, e_ids) and torch.equal(f_s, e_s)


def test_fused_exact_ties_fall_back(hnm_lib):
    """Duplicate item rows make exact ties at the cut: the certificate must refuse and the fallback decide by id."""
    from hnm_recommendation_b200.scorer import FusedScorer
    ue, ie = _emb(600, 2000, seed=9)
    ie[1000:] = ie[:1000]                      # every item has a twin 1000 ids later
    want_ids, want_s = O.recommend_exact(ue, ie, torch.arange(600), 12)
    sc = FusedScorer(ue.cuda(), ie.cuda())
    ids, s = sc.topk(None, 12)
    assert torch.equal(ids.cpu(), want_ids) and torch.equal(s.cpu(), want_s)
    # top-12 = 6 twins pairs -> the 12th and 13th never tie, but 11th/12th style ties inside are ordered by id
    assert (ids[:, 0] + 1000 == ids[:, 1]).all()


def test_lightgcn_recommend_uses_fused_path(hnm_lib):
    from hnm_recommendation_b200 import LightGCN, synth
    data = synth.interactions(3000, 1500, 40000, seed=3)
    m = LightGCN(data.num_users, data.num_items).to("cuda")
    with torch.no_grad():
        m.embeddings.weight.copy_(synth.trained_like_embeddings(4500, 64, seed=3))
    m.set_graph(data.edge_index())
    ue, ie = m.forward()
    uids = torch.arange(0, 3000, 7)
    got = m.recommend(uids)
    assert m._scorer is not None
    want, _ = O.recommend_exact(ue.cpu(), ie.cpu(), uids, 12)
    assert torch.equal(got.cpu(), want)
    allu = m.recommend_all()
    assert torch.equal(allu[uids.cuda()].cpu(), want)
    # filter = the user's own purchases (serve.py:350-352 semantics)
    hist = {}
    for u_, i_ in zip(data.users[:5000].tolist(), data.items[:5000].tolist()):
        hist.setdefault(u_, set()).add(i_)
    got_f = m.recommend(uids, filter_items=hist)
    want_f, _ = O.recommend_exact(ue.cpu(), ie.cpu(), uids, 12, hist)
    assert torch.equal(got_f.cpu(), want_f)


# ------------------------------------------------------------------------------- NeuralCF
NCF = golden_files("ncf")


@pytest.mark.parametrize("path", NCF, ids=lambda p: p.split("ncf_")[-1][:-4])
def test_ncf_golden(hnm_lib, path):
    from hnm_recommendation_b200 import NeuralCF
    g = load_golden(path)
    m = NeuralCF(int(g["num_users"]), int(g["num_items"]), mf_dim=int(g["mf_dim"]),
                 mlp_dims=[int(x) for x in g["mlp_dims"]], top_k=int(g["top_k"]))
    state = {k[len("state."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("state.")}
    assert set(state) == set(m.state_dict().keys())            # checkpoint-key compatibility
    m.load_state_dict(state)
    m = m.to("cuda").eval()
    logits = m(torch.from_numpy(g["user_ids"]), torch.from_numpy(g["item_ids"]))
    assert_close(logits, g["logits"], rtol=1e-5, atol_scale=1e-6, what="logits")     # P4
    one = m(torch.from_numpy(g["user_ids"][:1]), torch.from_numpy(g["item_ids"][:1]))
    assert one.dim() == 0
    alls = m.predict_all_items(torch.from_numpy(g["all_user_ids"]))
    assert_close(alls, g["all_scores"], rtol=1e-5, atol_scale=1e-6, what="all scores")
    rec = m.recommend(torch.from_numpy(g["all_user_ids"]))
    s = torch.from_numpy(g["all_scores"]).double()
    got_s = torch.gather(s, 1, rec.cpu())
    want_s = torch.gather(s, 1, torch.from_numpy(g["topk_canonical"]))
    assert_close(got_s, want_s, rtol=1e-5, atol_scale=1e-5, what="recommended scores")


def test_ncf_candidates_default_arch(hnm_lib):
    from hnm_recommendation_b200 import NeuralCF
    torch.manual_seed(0)
    m = NeuralCF(2000, 900).to("cuda").eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
    cand = torch.randint(0, 900, (2000, 50), dtype=torch.int32)
    out = m.score_candidates(None, cand)
    orc = O.NeuralCFOracle(2000, 900, state={k: v.cpu() for k, v in m.state_dict().items()})
    uu = torch.arange(2000).repeat_interleave(50)
    want = orc.forward(uu, cand.view(-1).long()).view(2000, 50)
    assert_close(out, want, rtol=1e-5, atol_scale=1e-6, what="candidate logits")
    # weight update invalidates the cached layer-1 tables
    with torch.no_grad():
        m.mlp_layers[0].bias.add_(0.1)
    out2 = m.score_candidates(None, cand)
    assert not torch.allclose(out, out2)
    with pytest.raises(IndexError):
        m(torch.tensor([2000]), torch.tensor([0]))


def test_ncf_recommend_filter_and_errors(hnm_lib):
    from hnm_recommendation_b200 import NeuralCF
    torch.manual_seed(1)
    m = NeuralCF(300, 120, mf_dim=16, mlp_dims=[32, 16, 8], top_k=7).to("cuda")
    uids = torch.tensor([3, 14, 15])
    s = m.predict_all_items(uids).cpu()
    filt = {3: set(torch.sort(s[0], descending=True).indices[:4].tolist()), 15: {0, 1}}
    rec = m.recommend(uids, filter_items=filt).cpu()
    s_f = s.clone()
    for r, u in enumerate(uids.tolist()):
        if u in filt:
            s_f[r, list(filt[u])] = float("-inf")
    assert torch.equal(rec, O.topk_canonical(s_f, 7))
    assert not m.training
    with pytest.raises(RuntimeError, match="out of range"):
        m.recommend(uids, k=121)
    with pytest.raises(IndexError):
        m.predict_all_items(torch.tensor([300]))
    with pytest.raises(RuntimeError):
        m(torch.tensor([1, 2]), torch.tensor([1]))
    cpu_model = NeuralCF(10, 10).eval()                       # scoring (eval / no_grad) has no CPU path
    with pytest.raises(RuntimeError, match="CUDA"):
        cpu_model(torch.tensor([1]), torch.tensor([1]))
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        NeuralCF(10, 10)(torch.tensor([1]), torch.tensor([1]))


def test_ncf_training_step_matches_reference_formulation(hnm_lib):
    """neural_cf.py:210-233: BCE-with-logits on the autograd forward (train mode); its eval-mode logits equal the
    fused kernels' (same parameters), gradients reach every parameter, and one optimiser step changes what the
    kernels score (the layer-1 tables follow the parameter versions)."""
    from hnm_recommendation_b200 import NeuralCF
    torch.manual_seed(5)
    m = NeuralCF(500, 200).to("cuda")
    u = torch.randint(0, 500, (256,), device="cuda")
    i = torch.randint(0, 200, (256,), device="cuda")
    labels = torch.randint(0, 2, (256,), device="cuda")
    m.eval()
    with torch.no_grad():
        k_logits = m(u, i)                                   # kernels
    ref_logits = m._forward_autograd(u, i)                   # reference formulation, dropout off in eval
    assert_close(k_logits, ref_logits.detach(), what="kernel vs autograd logits")
    m.train()
    loss = m.training_step({"user_ids": u, "item_ids": i, "labels": labels}, 0)
    assert loss.requires_grad and loss.dim() == 0
    loss.backward()
    assert all(p.grad is not None for p in m.parameters())
    opt = m.configure_optimizers()["optimizer"]
    opt.step()
    m.eval()
    with torch.no_grad():
        after = m(u, i)
    assert not torch.allclose(after, k_logits)
    assert_close(after, m._forward_autograd(u, i).detach(), what="kernel logits after an optimiser step")


def test_streamed_host_delivery_matches_device_result(hnm_lib):
    """topk(out_host=...): ids reach pinned host memory chunk by chunk on a copy stream; rows rewritten by the
    fallback tiers are patched.  Near-duplicate items force some users through the fallback."""
    from hnm_recommendation_b200.scorer import FusedScorer
    ue, ie = _emb(3000, 2000, seed=77)
    ie[1000:] = ie[:1000]                                     # exact ties: every user needs the fallback
    ie[1500:] += 1e-3                                          # ... except where the twins are told apart
    sc = FusedScorer(ue.cuda(), ie.cuda())
    want, _ = sc.topk(None, 12)
    assert sc.last_stats["uncertified"] > 0
    host = torch.empty(3000, 12, dtype=torch.int64).pin_memory()
    host.fill_(-1)
    got, _ = sc.topk(None, 12, out_host=host, chunk_users=640)          # five chunks
    assert torch.equal(got, want) and torch.equal(host, want.cpu())
    with pytest.raises(ValueError):
        sc.topk(None, 12, out_host=torch.empty(3000, 12, dtype=torch.int64))   # not pinned


def test_filter_own_history_stays_on_the_tensor_path(hnm_lib):
    """The reference's serving default (scripts/serve.py:350-352): every user's own purchases are filtered.  A
    user's purchases are also his best-scoring items, so without the exclusion signatures the nomination
    threshold would sit above every allowed item and most users would fall to the brute-force tier.  Lists must
    equal the exact kernel's with the same exclusion lists, bit for bit, and the fallback must stay rare."""
    from hnm_recommendation_b200 import LightGCN, engine, synth
    U, I, E = 30011, 8300, 700000
    data = synth.interactions(U, I, E, seed=11)
    m = LightGCN(U, I).to("cuda")
    with torch.no_grad():
        m.embeddings.weight.copy_(synth.trained_like_embeddings(U + I, 64, seed=11))
    m.set_graph(data.edge_index().cuda())
    ids, sc = m.recommend_all(return_scores=True, filter_purchased=True)
    stats = m._scorer.last_stats
    assert stats["tier3"] + stats["tier2"] <= U // 50, stats
    ptr, items = engine.history_csr(m.graph, U)
    assert int(ptr[-1]) == E and items.numel() == E                       # every interaction, repeats included
    ue, ie = m.forward()
    uids = torch.arange(0, U, 3, device="cuda")
    e_ids, e_sc = engine.topk_exact(ue, ie, uids, 12, engine.slice_csr(ptr, items, uids))
    assert torch.equal(ids[uids], e_ids) and torch.equal(sc[uids], e_sc)
    # nothing a user bought is recommended to him
    bought = torch.zeros(U, I, dtype=torch.bool, device="cuda")
    bought[torch.from_numpy(data.users).cuda(), torch.from_numpy(data.items).cuda()] = True
    assert not bool(torch.gather(bought, 1, ids).any())
    # the dict form of the reference API takes the same path for a batch of users
    some = uids[:300]
    hist = {int(u): set(items[int(ptr[u]):int(ptr[u + 1])].tolist()) for u in some.tolist()}
    assert torch.equal(m.recommend(some, filter_items=hist), ids[some])


def test_heavy_tailed_user_norms_stay_on_the_tensor_path(hnm_lib):
    """Per-user fp16 scaling: user rows with log-normal(sigma = 2) norms plus a few rows
    1e3 x larger.  With one scale per table those outliers push every other row toward fp16 subnormals and the
    users fall to the exact tier en masse; with one power of two per user row the ranking problem of every user is
    scaled on its own.  Bit-exact against the brute-force kernel, tier 3 <= 0.1 % of the users."""
    from hnm_recommendation_b200 import engine
    from hnm_recommendation_b200.scorer import FusedScorer
    U, I = 40000, 8300
    g = torch.Generator().manual_seed(21)
    ue = torch.randn(U, 64, generator=g) * 0.1
    ue *= torch.exp(2.0 * torch.randn(U, 1, generator=g))             # log-normal row norms, sigma = 2
    ue[torch.randint(0, U, (25,), generator=g)] *= 1e3                # a few outliers
    ue[7] = 0.0                                                       # and an all-zero row (scale 1, every score ties)
    ie = torch.randn(I, 64, generator=g) * 0.1
    ue, ie = ue.cuda(), ie.cuda()
    sc = FusedScorer(ue, ie)
    ids, s = sc.topk(None, 12)
    stats = sc.last_stats
    assert stats["tier3"] <= U // 1000, stats
    assert stats["uncertified"] <= U // 100, stats
    uids = torch.cat([torch.arange(0, U, 5), torch.tensor([7])]).cuda()
    w_ids, w_s = engine.topk_exact(ue, ie, uids, 12)
    assert torch.equal(ids[uids], w_ids) and torch.equal(s[uids], w_s)


@pytest.mark.parametrize("dim", [128, 256])
def test_fused_topk_wide_embeddings_bit_exact(hnm_lib, dim):
    """BASELINE.json configs[4] asks for d = 256: the tensor path takes the embedding dimension as 2 or 4 K chunks
    of 64 (user tiles resident in shared memory for the catalog sweep, B ring of (item tile, chunk) stages; the
    rescoring walks the slabs in order, so the fp64 chain still runs k = 0 .. d-1).  ids AND fp64 scores
    bit-identical to the oracle, with filters, on whole passes and on sliced left-over tiles."""
    from hnm_recommendation_b200 import engine
    from hnm_recommendation_b200.scorer import FusedScorer
    assert FusedScorer.supports(dim, 12, 5000)
    g = torch.Generator().manual_seed(dim)
    U, I = 3000, 5100
    ue = torch.randn(U, dim, generator=g) * 0.1
    ie = torch.randn(I, dim, generator=g) * 0.1 + 0.02
    sc = FusedScorer(ue.cuda(), ie.cuda())
    ids, s = sc.topk(None, 12)
    uids = torch.arange(0, U, 7)
    want_ids, want_s = O.recommend_exact(ue, ie, uids, 12)
    assert torch.equal(ids[uids.cuda()].cpu(), want_ids) and torch.equal(s[uids.cuda()].cpu(), want_s)
    w_ids, w_s = engine.topk_exact(ue.cuda(), ie.cuda(), None, 12)
    assert torch.equal(ids, w_ids) and torch.equal(s, w_s)
    assert sc.last_stats["tier3"] <= 3, sc.last_stats
    filt = {int(u): set(want_ids[r, :5].tolist()) for r, u in enumerate(uids[:40].tolist())}
    f_ids, f_s = sc.topk(uids[:40].cuda(), 12, filt)
    e_ids, e_s = O.recommend_exact(ue, ie, uids[:40], 12, filt)
    assert torch.equal(f_ids.cpu(), e_ids) and torch.equal(f_s.cpu(), e_s)
    # an item shard, k < 12, and more users than one pass of the persistent grid (whole passes + sliced tiles)
    U2 = (148 * 2 + 5) * 128 - 17
    ue2 = torch.randn(U2, dim, generator=g) * 0.1
    sh = FusedScorer(ue2.cuda(), ie[900:].cuda().contiguous(), item_begin=900)
    ids2, s2 = sh.topk(None, 7)
    chk = torch.cat([torch.arange(0, U2, 97), torch.arange(U2 - 700, U2)]).cuda()
    x_ids, x_s = engine.topk_exact(ue2.cuda(), ie[900:].cuda().contiguous(), chk, 7, item_begin=900)
    assert torch.equal(ids2[chk], x_ids) and torch.equal(s2[chk], x_s)


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [64, 128, 256])
def test_column_mean_matches_fp64_mean(hnm_lib, dim):
    """hnm_column_mean (the item centre): fp64 sums in a fixed order -> equal to torch's fp64 mean after the
    rounding to fp32, and identical from run to run."""
    from hnm_recommendation_b200._lib import call, ptr, stream
    torch.manual_seed(3)
    x = (torch.randn(70_001, dim, device="cuda") * 3 + 0.25).contiguous()
    ws_bytes = int(hnm_lib.hnm_column_mean_workspace_bytes(dim))
    ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device="cuda")
    outs = []
    for _ in range(2):
        out = torch.empty(dim, dtype=torch.float32, device="cuda")
        call("hnm_column_mean", ptr(x), x.size(0), dim, ptr(out), ptr(ws), ws_bytes, stream())
        outs.append(out)
    want = x.mean(dim=0, dtype=torch.float64)
    assert torch.equal(outs[0], outs[1])
    assert (outs[0].double() - want).abs().max().item() <= 2.0 ** -22 * want.abs().max().item() + 1e-12
