"""CPU-only checks: the C-ABI library builds, loads and exports every symbol the header declares;
the model mirrors keep the reference's API; host-side helpers."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol(hnm_lib):
    hdr = open(os.path.join(ROOT, "include", "hnm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hnm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 10
    missing = [n for n in sorted(declared) if not hasattr(hnm_lib, n)]
    assert not missing, f"declared in include/hnm_b200.h but not exported: {missing}"
    from hnm_recommendation_b200 import _lib
    assert declared == set(_lib._SIGNATURES), "ctypes signature table out of sync with the header"


def test_ctypes_signatures_match_header_arity():
    """Every prototype in include/hnm_b200.h has as many parameters as its ctypes argtypes entry."""
    from hnm_recommendation_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "hnm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"HNM_API\s+[\w\s\*]+?\b(hnm_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) == len(_lib._SIGNATURES)
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(_lib._SIGNATURES[name][1]), f"{name}: header has {n} parameters"


def test_abi_version_and_strerror(hnm_lib):
    assert hnm_lib.hnm_abi_version() == 2
    assert hnm_lib.hnm_strerror(0) == b"ok"
    assert b"NULL" in hnm_lib.hnm_strerror(-1)
    assert hnm_lib.hnm_strerror(-3) != hnm_lib.hnm_strerror(-2)
    assert hnm_lib.hnm_graph_build_workspace_bytes(0, 0, 0) == 0


def test_lightgcn_constructor_contract():
    # tests/test_models.py:162-171 of the reference
    from hnm_recommendation_b200 import LightGCN
    m = LightGCN(num_users=100, num_items=50, embedding_dim=16, top_k=5, num_layers=3)
    assert (m.num_users, m.num_items, m.num_layers, m.embedding_dim, m.top_k) == (100, 50, 3, 16, 5)
    assert len(m.alpha) == 4 and m.alpha == [0.25] * 4
    assert m.graph is None
    assert list(m.state_dict().keys()) == ["embeddings.weight"]
    assert tuple(m.embeddings.weight.shape) == (150, 16)
    bound = (6.0 / (150 + 16)) ** 0.5
    assert float(m.embeddings.weight.abs().max()) <= bound
    m2 = LightGCN(10, 5, alpha=0.5, num_layers=2)
    assert m2.alpha == pytest.approx([4 / 7, 2 / 7, 1 / 7])
    assert vars(m2.hparams)["alpha"] == 0.5 and vars(m2.hparams)["num_users"] == 10


def test_forward_before_set_graph_raises():
    from hnm_recommendation_b200 import LightGCN
    m = LightGCN(4, 3, embedding_dim=8)
    with pytest.raises(RuntimeError, match="Graph not set"):
        m.forward()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(hnm_lib):
    from hnm_recommendation_b200 import LightGCN
    m = LightGCN(4, 3, embedding_dim=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.set_graph(torch.tensor([[0, 4], [4, 0]]))


def test_exclusion_csr_host_logic():
    from hnm_recommendation_b200 import engine
    uids = torch.tensor([5, 2, 5, 9])
    ptr, items = engine.exclusion_csr(uids, {5: {7, 3}, 9: set(), 1: {4}}, "cpu")
    assert ptr.tolist() == [0, 2, 2, 4, 4] and items.tolist() == [3, 7, 3, 7]
    assert engine.exclusion_csr(uids, None, "cpu") == (None, None)
    assert engine.exclusion_csr(uids, {1: {2}}, "cpu") == (None, None)


def test_synth_is_deterministic_and_shaped():
    from hnm_recommendation_b200 import synth
    a = synth.interactions(1000, 300, 20000, seed=7)
    b = synth.interactions(1000, 300, 20000, seed=7)
    assert np.array_equal(a.users, b.users) and np.array_equal(a.items, b.items)
    assert a.users.shape == (20000,) and np.bincount(a.users, minlength=1000).min() >= 1
    assert a.items.min() >= 0 and a.items.max() < 300
    ei = a.edge_index()
    assert tuple(ei.shape) == (2, 40000) and int(ei[1, :20000].min()) >= 1000
    assert torch.equal(ei[0, :20000], ei[1, 20000:])


def test_metrics_standin():
    from hnm_recommendation_b200.metrics import RecommendationMetrics
    m = RecommendationMetrics(top_k=3)
    m.update(torch.tensor([[1, 2, 3], [4, 5, 6]]), [[1, 3], [9]])
    out = m.compute()
    assert set(out) == {"map_at_k", "recall_at_k", "precision_at_k", "ndcg_at_k"}
    assert out["recall_at_k"] == pytest.approx(0.5) and out["precision_at_k"] == pytest.approx(1 / 3)
    assert out["map_at_k"] == pytest.approx((1 + 2 / 3) / 2 / 2)


@pytest.mark.parametrize("user_tiles,item_tiles", [(1, 2), (8, 825), (63, 40), (444, 825), (1340, 825),
                                                   (10719, 825), (2680, 825), (445, 9), (7, 3)])
def test_fused_work_plan_covers_every_tile_pair_once(hnm_lib, user_tiles, item_tiles):
    """hnm_score_topk_fused_plan: whole-catalog passes of `mu` user tiles per CTA, the left-over tiles in groups of `mu`
    whose item range is sliced over the CTAs -- every (user tile, item tile) pair belongs to exactly one CTA."""
    import ctypes as C
    out = (C.c_int32 * 6)()
    assert hnm_lib.hnm_score_topk_fused_plan(user_tiles * 128, item_tiles * 128, out) == 0
    grid, full, tile0, triples, slices, mu = list(out)
    assert mu in (2, 3)
    assert grid >= 1 and tile0 == grid * mu * full and 0 <= user_tiles - tile0 < mu * grid
    assert triples == -(-(user_tiles - tile0) // mu)
    if triples:
        assert 1 <= slices <= grid
    seen = np.zeros((user_tiles, item_tiles), dtype=np.int32)
    for b in range(grid):
        for n in range(full):                       # CTA b, pass n: tiles (b*full + n)*mu .. +mu, all items
            t0 = (b * full + n) * mu
            seen[t0:t0 + mu, :] += 1
        for unit in range(b, triples * slices, grid):       # left-over units (group, slice): b, b + grid, ...
            j, sl = divmod(unit, slices)
            t0 = tile0 + mu * j
            i0, i1 = sl * item_tiles // slices, (sl + 1) * item_tiles // slices
            assert i1 > i0
            seen[t0:min(t0 + mu, user_tiles), i0:i1] += 1
    assert (seen == 1).all()
    if triples and slices > 1:
        # the tail is shorter than the unsliced one: rounds x slice length (seed tiles included) < one whole pass
        rounds = -(-triples * slices // grid)
        assert rounds * -(-item_tiles // slices) < item_tiles or rounds == 1
    ws = hnm_lib.hnm_score_topk_fused_workspace_bytes(user_tiles * 128, item_tiles * 128)
    need = triples * mu * 128 * slices * (256 * 20 + 12) if slices > 1 else 0    # 256 entries of 20 B + 2 counts + tau
    assert ws >= need and ws <= need + 4096
    assert hnm_lib.hnm_score_topk_fused_workspace_bytes(100, 128) < 0


def _cpu_graph(num_users, num_items, edges, seed, extra=None):
    """engine.Graph over CPU tensors built from the oracle's CSR (the host-side helpers are plain torch ops)."""
    import oracle as O
    from hnm_recommendation_b200 import engine
    gen = torch.Generator().manual_seed(seed)
    u = torch.randint(0, num_users, (edges,), generator=gen)
    i = torch.randint(0, num_items, (edges,), generator=gen) + num_users
    ei = torch.stack([torch.cat([u, i]), torch.cat([i, u])])
    if extra is not None:
        ei = torch.cat([ei, torch.tensor(extra, dtype=torch.long).t()], dim=1)
    n = num_users + num_items
    rowptr, col, _, dis = O.build_norm_adj(ei, None, n)
    return engine.Graph(n, int(col.numel()), rowptr.to(torch.int32), col.to(torch.int32), None, dis,
                        torch.zeros(0, dtype=torch.int32))


def test_user_shards_and_item_chunks_partition_every_item_row():
    """make_user_shard / make_item_chunks: for every item row the per-shard sub-ranges are contiguous, disjoint,
    ordered, and together cover exactly the row's user entries (everything but the self loop)."""
    from hnm_recommendation_b200 import engine
    U, I = 157, 41
    g = _cpu_graph(U, I, 900, seed=3)
    assert engine.is_bipartite(g, U)
    ic = engine.make_item_chunks(g, U, I, 64, num_chunks=5)
    assert ic is not None and len(ic.chunks) == 5
    assert [sh.u0 for sh in ic.chunks] == [0, 32, 64, 95, 126] and ic.chunks[-1].u1 == U
    rp, col = g.rowptr.long(), g.col.long()
    for r in range(I):
        lo, hi = int(rp[U + r]), int(rp[U + r + 1])
        assert int(col[hi - 1]) == U + r                                   # the self loop closes the row
        pos = lo
        for sh in ic.chunks:
            b, e = int(sh.seg_begin[r]), int(sh.seg_end[r])
            assert b == pos and b <= e
            assert bool(((col[b:e] >= sh.u0) & (col[b:e] < sh.u1)).all())
            pos = e
        assert pos == hi - 1
    assert engine.make_item_chunks(g, U, I, 64) is None                    # opt-in: nothing by default
    assert engine.make_item_chunks(g, U, I, 64, num_chunks=1) is None
    # heavy-row classification follows the SUB-range length
    sh = engine.make_user_shard(g, U, I, 0, U)
    lengths = (sh.seg_end - sh.seg_begin).long()
    assert sh.heavy_rows.numel() == int((lengths > g.heavy_threshold).sum()) == 0


def test_bipartite_check_sees_user_user_and_item_item_edges():
    from hnm_recommendation_b200 import engine
    U, I = 30, 20
    assert engine.is_bipartite(_cpu_graph(U, I, 200, seed=1), U)
    assert not engine.is_bipartite(_cpu_graph(U, I, 200, seed=1, extra=[(3, 7)]), U)              # user -> user
    assert not engine.is_bipartite(_cpu_graph(U, I, 200, seed=1, extra=[(U + 2, U + 5)]), U)      # item -> item
    g = _cpu_graph(U, I, 200, seed=1, extra=[(3, 7)])
    assert engine.make_item_chunks(g, U, I, 64, num_chunks=4) is None


def test_bench_gpu_comparators_run_device_agnostic():
    """bench.gpu_comparators (the eager-PyTorch restatement timed beside the kernels) is plain torch:
    it must run on CPU tensors too, report both matmul modes, and never raise."""
    import sys
    from types import SimpleNamespace
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    import oracle as O
    U, I = 200, 90
    g = _cpu_graph(U, I, 1500, seed=4)
    model = SimpleNamespace(graph=g, num_users=U, num_items=I, num_layers=3, alpha=O.layer_weights(3),
                            embeddings=SimpleNamespace(weight=torch.randn(U + I, 64) * 0.1))
    out = bench.gpu_comparators(model, sample_users=128, chunk=64)
    assert "error" not in out, out
    assert out["sample_users"] == 128 and out["propagate_ms"] > 0
    assert out["users_per_s_fp32"] > 0 and out["users_per_s_tf32"] > 0
    broken = SimpleNamespace(graph=None)
    assert "error" in bench.gpu_comparators(broken)


@pytest.mark.timeout(180)
def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU oracle timed on the host cores) prints ONE JSON line with the keys the
    driver reads; under torchrun only rank 0 works."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "config1",
           "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=170, env=dict(os.environ, RANK="0"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "users/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1"))
    assert other.returncode == 0 and not other.stdout.strip()


def test_header_is_c99_and_library_links_from_c(hnm_lib, tmp_path):
    """include/hnm_b200.h compiled by gcc as strict C99, linked against libhnm_b200.so, run without a GPU."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    libdir = os.path.join(ROOT, "hnm_recommendation_b200")
    exe = str(tmp_path / "abi_smoke")
    build = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                            os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe, "-L", libdir, "-l:libhnm_b200.so",
                            f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0, (run.returncode, run.stderr)
    assert run.stdout.strip().startswith("ok|")


def test_interaction_data_contract_roundtrip(tmp_path):
    """data.InteractionData: processed/train.parquet (customer_idx, article_idx; scripts/serve.py:174-177) ->
    get_graph() in the layout set_graph takes (scripts/train.py:221), the serving filter dict and its CSR form."""
    from hnm_recommendation_b200 import synth
    from hnm_recommendation_b200.data import InteractionData
    d = InteractionData.synthetic(300, 120, 2500, seed=3)
    path = d.to_parquet(str(tmp_path))
    assert path.endswith(os.path.join("processed", "train.parquet"))
    back = InteractionData.from_parquet(str(tmp_path), num_users=300, num_items=120)
    ei, ew = back.get_graph()
    assert ew is None and torch.equal(ei, synth.interactions(300, 120, 2500, seed=3).edge_index())
    assert ei.dtype == torch.int64 and int(ei[0, :2500].max()) < 300 and int(ei[1, :2500].min()) >= 300
    hist = back.user_history()
    assert set(hist) == set(range(300))                                   # every synthetic user buys at least once
    ptr, items = back.history_csr()
    assert ptr.numel() == 301 and int(ptr[-1]) == items.numel() == sum(len(v) for v in hist.values())
    for u in (0, 17, 299):
        row = items[int(ptr[u]):int(ptr[u + 1])].tolist()
        assert row == sorted(hist[u])
    with pytest.raises(ValueError):
        InteractionData.from_arrays([0, 5], [1, 2], num_users=3)


def test_graft_entry_build_runs_here():
    """The driver's "does it build" check: build() compiles (incrementally) for sm_100a, loads the library and
    checks its ABI version against the host side's."""
    import __graft_entry__ as entry
    entry.build()


def test_entry_points_validate_arguments_before_touching_the_device(hnm_lib):
    """The error contract of include/hnm_b200.h (0 / negative argument error / positive cudaError_t): argument errors
    are reported before any CUDA call, so they can be checked here without a GPU.  The pointers are never read."""
    import ctypes as C
    L = hnm_lib
    E_NULL, E_RANGE, E_DIM, E_ALIGN = -1, -2, -3, -5
    p, odd = C.c_void_p(0x10000), C.c_void_p(0x10008)            # non-NULL; `odd` is 8- but not 16-byte aligned
    N = None
    # LightGCN.predict (src/models/lightgcn.py:166-186)
    assert L.hnm_pair_scores(N, p, p, p, 4, 64, 10, 10, p, N, N) == E_NULL
    assert L.hnm_pair_scores(p, p, p, p, 0, 64, 10, 10, p, N, N) == 0                       # empty batch: nothing to do
    assert L.hnm_pair_scores(p, p, p, p, -1, 64, 10, 10, p, N, N) == E_RANGE
    assert L.hnm_pair_scores(p, p, p, p, 4, 0, 10, 10, p, N, N) == E_RANGE
    # predict_all_items (:188-204)
    assert L.hnm_score_all_items(p, p, N, 0, 100, 64, p, N) == 0
    assert L.hnm_score_all_items(p, N, N, 4, 100, 64, p, N) == E_NULL
    assert L.hnm_score_all_items(p, p, N, 4, 100, 0, p, N) == E_RANGE
    # exact top-k (:349-356): exclusion CSR needs both arrays; k within [1, items]
    assert L.hnm_topk_exact(p, p, N, 4, 0, 100, 64, 12, p, N, 1, p, p, N) == E_NULL
    assert L.hnm_topk_exact(p, p, N, 4, 0, 100, 64, 0, N, N, 1, p, p, N) == E_RANGE
    assert L.hnm_topk_exact(p, p, N, 4, 0, 8, 64, 12, N, N, 1, p, p, N) == E_RANGE          # torch.topk: k out of range
    assert L.hnm_topk_exact(p, p, N, 4, 5, 5, 64, 1, N, N, 1, p, p, N) == E_RANGE           # empty item shard
    assert L.hnm_topk_exact(p, p, N, 0, 0, 100, 64, 12, N, N, 1, p, p, N) == 0
    # merge of per-shard lists (multi-GPU item shards)
    assert L.hnm_merge_topk(p, p, 0, 4, 12, p, p, N) == E_RANGE
    assert L.hnm_merge_topk(p, p, 65, 4, 12, p, p, N) == E_RANGE
    assert L.hnm_merge_topk(p, N, 2, 4, 12, p, p, N) == E_NULL
    assert L.hnm_merge_topk(p, p, 2, 0, 12, p, p, N) == 0
    # dense select (NeuralCF.recommend, src/models/neural_cf.py:300-326)
    assert L.hnm_topk_dense(p, 4, 10, N, N, 12, p, N, N) == E_RANGE
    assert L.hnm_topk_dense(p, 4, 100, p, N, 12, p, N, N) == E_NULL
    assert L.hnm_topk_dense(N, 4, 100, N, N, 12, p, N, N) == E_NULL
    # fused score/select: padded sizes are whole tiles, operands 128-byte aligned
    a = [p, 1000, 1024, p, 500, 512, 64, 15, p, 192, p, p, N, N, 0, N]
    assert L.hnm_score_topk_fused(*a) not in (E_NULL, E_RANGE, E_ALIGN)                     # arguments fine: fails later, on the device check
    bad = list(a); bad[2] = 1000
    assert L.hnm_score_topk_fused(*bad) == E_RANGE
    bad = list(a); bad[7] = 33
    assert L.hnm_score_topk_fused(*bad) == E_RANGE                                           # kth_sel beyond the 32 buckets
    bad = list(a); bad[9] = 191
    assert L.hnm_score_topk_fused(*bad) == E_RANGE                                           # odd list capacity
    bad = list(a); bad[0] = odd
    assert L.hnm_score_topk_fused(*bad) == E_ALIGN
    bad = list(a); bad[8] = N
    assert L.hnm_score_topk_fused(*bad) == E_NULL
    assert L.hnm_score_topk_fused_plan(1024, 512, N) == E_NULL
    assert L.hnm_score_topk_fused_workspace_bytes(1000, 512) < 0
    # exact rescoring: dimension, k, capacity, alignment
    r = [p, p, N, 4, 64, 0, 500, p, 192, p, p, p, p, N, N, N, 12, p, p, p, N]
    bad = list(r); bad[4] = 96
    assert L.hnm_rescore_topk(*bad) == E_DIM
    bad = list(r); bad[16] = 33
    assert L.hnm_rescore_topk(*bad) == E_RANGE
    bad = list(r); bad[8] = 193
    assert L.hnm_rescore_topk(*bad) == E_RANGE
    bad = list(r); bad[1] = odd
    assert L.hnm_rescore_topk(*bad) == E_ALIGN
    bad = list(r); bad[14] = p                                                               # excl_ptr without excl_items
    assert L.hnm_rescore_topk(*bad) == E_NULL
    bad = list(r); bad[3] = 0
    assert L.hnm_rescore_topk(*bad) == 0
    # propagation: a layer cannot run in place (rows are gathered while others are written)
    assert L.hnm_lightgcn_layer(p, p, N, p, p, p, p, 0.25, 100, 64, 0, 100, N, 0, 0, 1024, 0, N) == E_RANGE
    # set_graph (:81-112): outputs and workspace required, edge weights and CSR weights go together, sizes sane,
    # workspace at least hnm_graph_build_workspace_bytes
    nh = (C.c_int32 * 1)()
    assert L.hnm_graph_build(p, p, N, 10, 5, N, p, N, p, 1024, p, nh, p, 1 << 20, N) == E_NULL
    assert L.hnm_graph_build(N, p, N, 10, 5, p, p, N, p, 1024, p, nh, p, 1 << 20, N) == E_NULL
    assert L.hnm_graph_build(p, p, p, 10, 5, p, p, N, p, 1024, p, nh, p, 1 << 20, N) == E_NULL     # edge_w without csr_w
    assert L.hnm_graph_build(p, p, N, 10, 0, p, p, N, p, 1024, p, nh, p, 1 << 20, N) == E_RANGE
    assert L.hnm_graph_build(p, p, N, 2 ** 31, 5, p, p, N, p, 1024, p, nh, p, 1 << 20, N) == E_RANGE  # nnz beyond int32
    need = L.hnm_graph_build_workspace_bytes(5, 10, 0)
    assert need > 0 and L.hnm_graph_build(p, p, N, 10, 5, p, p, N, p, 1024, p, nh, p, need - 1, N) == -4
    # NeuralCF (src/models/neural_cf.py:112-141,143-208)
    assert L.hnm_ncf_score_pairs(p, p, p, p, p, p, 2, p, 0.0, N, p, 4, 64, p, N) == E_NULL
    assert L.hnm_ncf_score_candidates(p, p, p, p, p, p, 2, p, 0.0, p, 4, p, 0, 64, p, N) == E_RANGE
    # every code has its own message
    msgs = {c: L.hnm_strerror(c).decode() for c in (0, -1, -2, -3, -4, -5, -6, -7)}
    assert all(msgs.values()) and len(set(msgs.values())) == len(msgs)
