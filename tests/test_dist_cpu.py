"""world_size-2 gloo test of the multi-GPU host logic (partitioning, the three exchanges, merge order).
The kernels are replaced by the oracle through the Backend hook, so this runs without a GPU."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleBackend:
    """Same contract as hnm_recommendation_b200.dist.CudaBackend, computed on CPU by the oracle."""

    def propagate(self, model, my_ranges, exchange, exchange_final):
        rowptr, col, val, _ = model.graph
        w = model.embeddings.weight
        acc = model.alpha[0] * w
        cur = w
        for layer in range(1, model.num_layers + 1):
            full = O.lightgcn_oracle._spmm(rowptr, col, val, cur)
            nxt = torch.full_like(w, float("nan"))            # rows this rank does not own stay garbage
            for r0, r1 in my_ranges:
                nxt[r0:r1] = full[r0:r1]
                acc[r0:r1] += model.alpha[layer] * full[r0:r1]
            if layer < model.num_layers:
                exchange(nxt)
                assert not torch.isnan(nxt).any(), "exchange left rows unfilled"
            cur = nxt
        exchange_final(acc)
        return acc

    def local_topk(self, sharded, ue, ie_shard, item_begin, k, defer=False):
        ids, sc = O.recommend_exact(ue, ie_shard.contiguous(), torch.arange(ue.size(0)), k)
        return ids + item_begin, sc

    def pending_flag(self, sharded, device):
        return torch.zeros(1, dtype=torch.int64, device=device)

    def finalize_topk(self, sharded):
        return False

    def merge(self, ids, scores):
        g, n, k = ids.shape
        ids = ids.permute(1, 0, 2).reshape(n, g * k)
        sc = scores.permute(1, 0, 2).reshape(n, g * k)
        o1 = torch.sort(ids, dim=1, stable=True).indices                       # id ascending ...
        sc1, ids1 = torch.gather(sc, 1, o1), torch.gather(ids, 1, o1)
        o2 = torch.sort(sc1, dim=1, descending=True, stable=True).indices      # ... then score descending, stable
        return torch.gather(ids1, 1, o2)[:, :k].contiguous(), torch.gather(sc1, 1, o2)[:, :k].contiguous()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hnm_recommendation_b200 import dist as hdist
        torch.manual_seed(0)
        U, I, d, L, k = 203, 97, 16, 3, 12                     # sizes that do not divide by the world size
        g = torch.Generator().manual_seed(1)
        u = torch.randint(0, U, (1500,), generator=g)
        i = torch.randint(0, I, (1500,), generator=g) + U
        ei = torch.stack([torch.cat([u, i]), torch.cat([i, u])])
        w = torch.randn(U + I, d, generator=g) * 0.1
        model = SimpleNamespace(num_users=U, num_items=I, top_k=k, num_layers=L, alpha=O.layer_weights(L),
                                graph=O.build_norm_adj(ei, None, U + I), embeddings=SimpleNamespace(weight=w))
        sh = hdist.ShardedLightGCN(model, backend=OracleBackend(), mode="items")
        plan = sh.plan
        assert plan.user_rows[0][0] == 0 and plan.user_rows[-1][1] == U
        assert plan.item_rows[0][0] == U and plan.item_rows[-1][1] == U + I
        assert sum(b - a for a, b in plan.item_shards) == I
        ue, ie = sh.forward()
        rowptr, col, val, _ = model.graph
        ou, oi = O.forward(w, rowptr, col, val, U, L, model.alpha)
        assert torch.allclose(ue, ou, rtol=1e-6, atol=1e-8) and torch.allclose(ie, oi, rtol=1e-6, atol=1e-8)
        ids, sc = sh.recommend_all(return_scores=True)
        want_ids, want_sc = O.recommend_exact(ou, oi, torch.arange(U), k)
        ok = torch.equal(ids, want_ids) and torch.allclose(sc, want_sc, rtol=1e-12, atol=0)
        # default mode: users are sharded, nothing but the results is exchanged
        sh_u = hdist.ShardedLightGCN(model, backend=OracleBackend(), mode="users")
        ids_u, sc_u = sh_u.recommend_all(return_scores=True)
        ok = ok and torch.equal(ids_u, want_ids) and torch.allclose(sc_u, want_sc, rtol=1e-12, atol=0)
        ue_u, ie_u = sh_u.forward(all_rows=False)
        a, b = sh_u.plan.user_rows[rank]
        ok = ok and torch.allclose(ue_u[a:b], ou[a:b], rtol=1e-6, atol=1e-8) and torch.allclose(ie_u, oi, rtol=1e-6, atol=1e-8)
        # a new table from the host: a rank uploads only the rows it owns, the item block is completed by exchange
        w_old = w.clone()
        w2 = torch.randn(U + I, d, generator=torch.Generator().manual_seed(5))
        nbytes = sh_u.load_embeddings_from_host(w2)
        i0, i1 = sh_u.plan.item_rows[rank]
        ok = ok and nbytes == ((b - a) + (i1 - i0)) * d * 4
        ok = ok and torch.equal(w[a:b], w2[a:b]) and torch.equal(w[U:], w2[U:])
        other = [r for r in range(world) if r != rank][0]
        oa, ob = sh_u.plan.user_rows[other]
        ok = ok and torch.equal(w[oa:ob], w_old[oa:ob])                 # nobody sent what this rank does not read
        sh.load_embeddings_from_host(w2)                                # "items" mode reads every row
        ok = ok and torch.equal(w, w2)
        # item shards smaller than k: every shard hands in what it has, sentinels fill the rest of its list
        U2, I2, k2 = 37, 9, 6                                           # shards of 5 + 4 items (world 2) or 3 + 3 + 3 (world 3)
        w3 = torch.randn(U2 + I2, d, generator=torch.Generator().manual_seed(7)) * 0.1
        u3 = torch.randint(0, U2, (200,), generator=g)
        i3 = torch.randint(0, I2, (200,), generator=g) + U2
        ei3 = torch.stack([torch.cat([u3, i3]), torch.cat([i3, u3])])
        small = SimpleNamespace(num_users=U2, num_items=I2, top_k=k2, num_layers=L, alpha=O.layer_weights(L),
                                graph=O.build_norm_adj(ei3, None, U2 + I2), embeddings=SimpleNamespace(weight=w3))
        sh_s = hdist.ShardedLightGCN(small, backend=OracleBackend(), mode="items")
        ids_s, sc_s = sh_s.recommend_all(return_scores=True)
        rp3, c3, v3, _ = small.graph
        ou3, oi3 = O.forward(w3, rp3, c3, v3, U2, L, small.alpha)
        want_i3, want_s3 = O.recommend_exact(ou3, oi3, torch.arange(U2), k2)
        ok = ok and torch.equal(ids_s, want_i3) and torch.allclose(sc_s, want_s3, rtol=1e-12, atol=0)
        q.put((rank, bool(ok), ""))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lightgcn_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, ok, err in results:
        assert ok, f"rank {rank} failed:\n{err}"


def test_even_ranges_and_plan():
    from hnm_recommendation_b200.dist import ShardPlan, even_ranges
    assert even_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert even_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    from hnm_recommendation_b200.dist import chunk_ranges
    assert chunk_ranges(10, 3) == [(0, 4), (4, 8), (8, 10)]
    assert chunk_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert chunk_ranges(9, 4) == [(0, 3), (3, 6), (6, 9), (9, 9)]
    assert chunk_ranges(1371980, 8)[7] == (7 * 171498, 1371980)
    p = ShardPlan.make(1371980, 105542, 8, 3)
    assert p.user_rows[7][1] == 1371980 and p.item_rows[0][0] == 1371980 and p.item_rows[7][1] == 1371980 + 105542
    sizes = [b - a for a, b in p.item_shards]
    assert max(sizes) - min(sizes) <= 1
