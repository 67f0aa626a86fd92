"""Multi-GPU form of the hot path: one process per GPU, torch.distributed for the plumbing.

No counterpart in the reference (single device, scripts/train.py:233-234).  BASELINE.json's north_star
prescribes item-catalog shards + allgather + merge for the scoring and row shards + allgather for the
propagation; both are implemented (mode "items", the row-sliced propagation), but the DEFAULT is the
user partitioning below, which moves 13x fewer bytes per layer and lets every stage shrink as 1/G
(the north_star form is timed beside it in every multi-GPU bench line as items_mode_ms):

  * propagation: the USERS are partitioned.  A rank computes its own user rows (they gather from the
    item block, which every rank holds) and, for every item row, the partial neighbour sum over its own
    users; the partial sums (27 MB at the H&M shape -- an all-gather of the user block would move 351 MB)
    are exchanged inside the kernels over NVLink peer memory -- stored straight into the owning rank's
    staging buffer, reduced there in rank order and written back into every rank's table
    (engine.propagate_user_sharded_peer) -- or, with HNM_PEER_EXCHANGE=0 / on gloo, by one all-reduce per
    layer followed by a local finish pass;
  * scoring, mode "items" (north_star): the item catalog is sharded; each rank runs the fused
    score/select + exact rescoring against its item shard for ALL users, the per-shard exact top-k
    lists are exchanged (all-to-all by user slice), merged by (score desc, item id asc), and the
    merged slices are all-gathered;
  * scoring, mode "users" (default): users are independent units, so each rank scores the users of
    its own row slice against the whole catalog (27 MB, replicated) with no data-path collective;
    only the [U/G, k] results are all-gathered, and the last propagation exchange shrinks to the
    item rows.  Per-rank work (nomination, rescoring, fallback) then falls as 1/G, which the item
    mode cannot do: every item shard nominates ~k candidates for every user.

The compute steps are injected (``Backend``) so the host logic -- partitioning, exchanges, merge
order -- can be exercised on CPU with gloo and the oracle standing in for the kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def even_ranges(n: int, parts: int) -> List[Tuple[int, int]]:
    """Split [0, n) into `parts` contiguous ranges whose sizes differ by at most one."""
    base, rem = divmod(n, parts)
    out, start = [], 0
    for p in range(parts):
        size = base + (1 if p < rem else 0)
        out.append((start, start + size))
        start += size
    return out


def chunk_ranges(n: int, parts: int) -> List[Tuple[int, int]]:
    """Split [0, n) into `parts` contiguous ranges of ceil(n / parts) rows, the last one(s) shorter: slices of one
    size laid out back to back, so that ONE in-place all-gather assembles [0, n) with the padding at the very end."""
    size = -(-n // parts) if n > 0 else 0
    return [(min(n, p * size), min(n, (p + 1) * size)) for p in range(parts)]


@dataclass
class ShardPlan:
    world: int
    rank: int
    user_rows: List[Tuple[int, int]]      # per rank, node-id range inside [0, U)
    item_rows: List[Tuple[int, int]]      # per rank, node-id range inside [U, U+I)
    item_shards: List[Tuple[int, int]]    # per rank, item-index range inside [0, I)
    user_slices: List[Tuple[int, int]]    # per rank, the users whose lists this rank merges

    @staticmethod
    def make(num_users: int, num_items: int, world: int, rank: int) -> "ShardPlan":
        ur = chunk_ranges(num_users, world)     # users: equal chunks (the result all-gather needs no compaction)
        ir = even_ranges(num_items, world)
        return ShardPlan(world, rank, ur, [(num_users + a, num_users + b) for a, b in ir], ir, ur)


class Collectives:
    """The three exchanges, with NCCL fast paths and a plain fallback that also runs on gloo."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nccl = dist.get_backend(group) == "nccl"

    def allgather_rows(self, buf: torch.Tensor, ranges: Sequence[Tuple[int, int]]) -> None:
        """Every rank has written buf[ranges[rank]]; make all ranges valid on all ranks (in place)."""
        views = [buf[a:b] for a, b in ranges]
        sizes = {v.shape[0] for v in views}
        contiguous = all(ranges[i][1] == ranges[i + 1][0] for i in range(len(ranges) - 1))
        if len(sizes) == 1 and contiguous and self.world > 1:
            # equal slices laid out back to back: one in-place all-gather, no staging copies
            dist.all_gather_into_tensor(buf[ranges[0][0]:ranges[-1][1]], views[self.rank], group=self.group)
        elif self.nccl:
            # uneven slices: every rank sends its slice to every peer, coalesced into one NCCL group
            ops = []
            for peer in range(self.world):
                if peer == self.rank:
                    continue
                gp = dist.get_global_rank(self.group, peer) if self.group else peer
                if views[self.rank].numel():
                    ops.append(dist.P2POp(dist.isend, views[self.rank], gp, group=self.group))
                if views[peer].numel():
                    ops.append(dist.P2POp(dist.irecv, views[peer], gp, group=self.group))
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        else:
            for src, v in enumerate(views):
                if v.numel():
                    dist.broadcast(v, src=dist.get_global_rank(self.group, src) if self.group else src,
                                   group=self.group)

    def exchange_by_user_slice(self, t: torch.Tensor, slices: Sequence[Tuple[int, int]]) -> torch.Tensor:
        """t: [U, k] computed against this rank's item shard.  Returns [world, n_mine, k]: every rank's
        rows for the users of my slice."""
        a, b = slices[self.rank]
        n_mine = b - a
        out = torch.empty((self.world, n_mine) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        if self.nccl:
            in_splits = [y - x for x, y in slices]
            dist.all_to_all_single(out.view(self.world * n_mine, *t.shape[1:]), t.contiguous(),
                                   output_split_sizes=[n_mine] * self.world, input_split_sizes=in_splits,
                                   group=self.group)
        else:
            full = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(full, t.contiguous(), group=self.group)
            for r in range(self.world):
                out[r] = full[r][a:b]
        return out

    def allgather_slices(self, mine: torch.Tensor, slices: Sequence[Tuple[int, int]], total: int,
                         dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """Every rank holds the rows `slices[rank]` of a [total, ...] result: return all of it on every rank,
        optionally converted to `dtype` on the way in (int64 item ids travel as int32).

        Slices of one size with only the tail shorter (chunk_ranges) take ONE in-place all-gather into a buffer of
        world x size rows whose first `total` rows ARE the result: no staging copy of the input (the conversion
        writes it in place) and no compaction afterwards (round 2 first padded uneven slices and dropped the
        padding with an index_select: 0.18 ms per 16 MB, a tenth of the 8-GPU step)."""
        dtype = dtype or mine.dtype
        sizes = [b - a for a, b in slices]
        rows = sizes[0] if sizes else 0
        starts_ok = all(slices[r][0] == min(total, r * rows) for r in range(self.world))
        if self.world > 1 and rows > 0 and starts_ok and all(sz <= rows for sz in sizes):
            flat = torch.empty((self.world * rows,) + tuple(mine.shape[1:]), dtype=dtype, device=mine.device)
            lo = self.rank * rows
            flat[lo:lo + sizes[self.rank]].copy_(mine)
            src = flat[lo:lo + rows]
            dist.all_gather_into_tensor(flat, src if self.nccl else src.clone(), group=self.group)
            return flat[:total]
        out = torch.empty((total,) + tuple(mine.shape[1:]), dtype=dtype, device=mine.device)
        a, b = slices[self.rank]
        out[a:b] = mine
        self.allgather_rows(out, slices)
        return out


class ShardedLightGCN:
    """Wraps a LightGCN whose parameters and graph are replicated on every rank."""

    def __init__(self, model, group=None, backend=None, mode: Optional[str] = None):
        import os
        self.model = model
        self.mode = mode or os.environ.get("HNM_SHARD_MODE", "users")
        if self.mode not in ("users", "items"):
            raise ValueError("mode must be 'users' or 'items'")
        self.coll = Collectives(group)
        self.plan = ShardPlan.make(model.num_users, model.num_items, self.coll.world, self.coll.rank)
        self.backend = backend or CudaBackend()
        self._scorer = None
        self.stage_ms: Dict[str, float] = {}

    # ---------------------------------------------------------------- propagate
    def forward(self, all_rows: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """(user_embeddings, item_embeddings).  With all_rows=False in "users" mode only this rank's
        user rows (plan.user_rows[rank]) and all item rows are valid."""
        m, p = self.model, self.plan
        if m.graph is None:
            raise RuntimeError("Graph not set. Call set_graph() first.")
        my_ranges = [p.user_rows[p.rank], p.item_rows[p.rank]]

        def exchange(buf):
            self.coll.allgather_rows(buf, p.user_rows)
            self.coll.allgather_rows(buf, p.item_rows)

        def exchange_items(buf):
            self.coll.allgather_rows(buf, p.item_rows)

        if hasattr(self.backend, "propagate_sharded") and self.backend.can_partition_users(m):
            return self.backend.propagate_sharded(self, all_rows or self.mode != "users")
        final = self.backend.propagate(m, my_ranges, exchange,
                                       exchange_items if (self.mode == "users" and not all_rows) else exchange)
        self._scorer = None
        return final[: m.num_users], final[m.num_users:]

    # ---------------------------------------------------------------- parameters from the host
    def load_embeddings_from_host(self, table: torch.Tensor) -> int:
        """Upload a new embedding table ([U + I, d] fp32, pinned host memory, the same on every rank) so that it
        crosses PCIe ONCE over the whole job: every rank copies the rows it owns -- its users and its 1/G slice of
        the item block -- and the item block (all of which every rank reads) is completed over NVLink with one
        in-place all-gather.  In "items" mode the user rows are all-gathered too.  Enqueued on the current stream,
        no host synchronisation.  Returns the host-to-device bytes of this rank."""
        m, p = self.model, self.plan
        w = m.embeddings.weight.data
        if tuple(table.shape) != tuple(w.shape) or table.dtype != w.dtype:
            raise ValueError("table must have the shape and dtype of embeddings.weight")
        rows = 0
        for a, b in (p.user_rows[p.rank], p.item_rows[p.rank]):
            if b > a:
                w[a:b].copy_(table[a:b], non_blocking=True)
                rows += b - a
        if self.coll.world > 1:
            self.coll.allgather_rows(w, p.item_rows)
            if self.mode != "users":
                self.coll.allgather_rows(w, p.user_rows)
        if hasattr(m, "invalidate"):
            m.invalidate()
        self._scorer = None
        return rows * w.size(1) * w.element_size()

    # ---------------------------------------------------------------- recommend
    def recommend_all(self, k: Optional[int] = None, return_scores: bool = False):
        m, p = self.model, self.plan
        k = m.top_k if k is None else int(k)
        with torch.no_grad():
            if self.mode == "users":
                ue, ie = self.forward(all_rows=False)
                u0, u1 = p.user_slices[p.rank]
                # my users vs all items; the scorer's single host read is deferred until the all-gather is enqueued
                ids, sc = self.backend.local_topk(self, ue[u0:u1], ie, 0, k, defer=True)

                def gather():
                    if m.num_items < 2 ** 31:
                        # item ids fit 32 bits: half the bytes on the wire (66 instead of 132 MB at the H&M shape)
                        return self.coll.allgather_slices(ids, p.user_slices, m.num_users, dtype=torch.int32).long()
                    return self.coll.allgather_slices(ids, p.user_slices, m.num_users)

                out_ids = gather()
                # every rank's "something is left for tier 3" flag travels behind the ids, so that all ranks take
                # the same decision from ONE host read (the step's only synchronisation) without another collective
                mine = self.backend.pending_flag(self, ids.device)
                flags = torch.empty(self.coll.world, dtype=torch.int64, device=ids.device)
                dist.all_gather_into_tensor(flags, mine, group=self.coll.group)
                anyone = bool(flags.sum().item())
                self.backend.finalize_topk(self)
                if anyone:                       # rare: some rank patched rows after the fact, gather again
                    out_ids = gather()
                if return_scores:
                    return out_ids, self.coll.allgather_slices(sc, p.user_slices, m.num_users)
                return out_ids
            ue, ie = self.forward()
            i0, i1 = p.item_shards[p.rank]
            if k > m.num_items:
                raise ValueError(f"top_k = {k} exceeds the catalog ({m.num_items} items)")
            # a shard smaller than k (tiny catalogs, many ranks) contributes all it has; the rest of its list is
            # (-inf, INT64_MAX) sentinels, which the merge's (score desc, id asc) order puts behind every item
            k_loc = min(k, i1 - i0)
            if k_loc > 0:
                ids, sc = self.backend.local_topk(self, ue, ie[i0:i1], i0, k_loc)   # [U, k_loc] vs my item shard
            else:
                ids = torch.empty(m.num_users, 0, dtype=torch.int64, device=ue.device)
                sc = torch.empty(m.num_users, 0, dtype=torch.float64, device=ue.device)
            if k_loc < k:
                ids = torch.cat([ids, ids.new_full((m.num_users, k - k_loc), torch.iinfo(torch.int64).max)], dim=1)
                sc = torch.cat([sc, sc.new_full((m.num_users, k - k_loc), float("-inf"))], dim=1)
            ids_x = self.coll.exchange_by_user_slice(ids, p.user_slices)           # [G, n_mine, k]
            sc_x = self.coll.exchange_by_user_slice(sc, p.user_slices)
            m_ids, m_sc = self.backend.merge(ids_x, sc_x)                          # [n_mine, k]
            out_ids = self.coll.allgather_slices(m_ids, p.user_slices, m.num_users)
            if return_scores:
                return out_ids, self.coll.allgather_slices(m_sc, p.user_slices, m.num_users)
        return out_ids


class CudaBackend:
    """The B200 kernels (default)."""

    def __init__(self):
        self._shard_key = None
        self._shard = None

    def can_partition_users(self, model) -> bool:
        """The user-partitioned form needs a bipartite graph (item rows hold users and their self loop only);
        anything else takes the row-sliced form with an all-gather per layer."""
        from . import engine
        return model.embedding_dim % 4 == 0 and engine.is_bipartite(model.graph, model.num_users)

    def propagate_sharded(self, sharded, final_users: bool):
        """Users partitioned over the ranks; per layer one all-reduce of the item block (engine.propagate_user_sharded)."""
        from . import engine
        m, coll, plan = sharded.model, sharded.coll, sharded.plan
        key = (id(m.graph), coll.world, coll.rank)
        if self._shard_key != key:
            if m.graph is None:
                raise RuntimeError("Graph not set. Call set_graph() first.")
            u0, u1 = plan.user_rows[plan.rank]
            self._shard = engine.make_user_shard(m.graph, m.num_users, m.num_items, u0, u1)
            self._shard_key = key

        def allreduce(buf):
            # asynchronous on NCCL: the caller overlaps it with the rank's own user rows and waits on the handle
            return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=coll.group, async_op=coll.nccl)

        pb = self._peer_buffers(m, coll)
        if pb is not None:
            acc = engine.propagate_user_sharded_peer(m.graph, self._shard, m.embeddings.weight, m.alpha, m.num_layers,
                                                     m.num_users, pb)
        else:
            acc = engine.propagate_user_sharded(m.graph, self._shard, m.embeddings.weight, m.alpha, m.num_layers,
                                                m.num_users, allreduce)
        if final_users:
            coll.allgather_rows(acc, plan.user_rows)
        sharded._scorer = None
        return acc[: m.num_users], acc[m.num_users:]

    def _peer_buffers(self, m, coll):
        """Peer-mapped exchange buffers (NCCL process groups on one node; HNM_PEER_EXCHANGE=0 keeps the NCCL
        all-reduce form).  Allocated once per (graph, world); None when symmetric memory is unavailable."""
        import os
        if not coll.nccl or os.environ.get("HNM_PEER_EXCHANGE", "1") == "0":
            return None
        key = (id(m.graph), coll.world, coll.rank, m.embedding_dim)
        if getattr(self, "_peer_key", None) != key:
            from . import engine
            try:
                self._peer = engine.PeerBuffers(m.num_nodes, m.num_items, m.embedding_dim, coll.group)
            except Exception as exc:  # noqa: BLE001
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({type(exc).__name__}: {exc}); using NCCL all-reduce")
                self._peer = None
            self._peer_key = key
        return self._peer

    def propagate(self, model, my_ranges, exchange, exchange_final):
        from . import engine
        return engine.propagate(model.graph, model.embeddings.weight, model.alpha, model.num_layers,
                                row_ranges=my_ranges, exchange=exchange, exchange_final=exchange_final)

    def local_topk(self, sharded, ue, ie_shard, item_begin, k, defer=False):
        from . import engine
        from .scorer import FusedScorer
        ie_shard = ie_shard.contiguous()
        ue = ue.contiguous()
        if FusedScorer.supports(ue.size(1), k, ie_shard.size(0)):
            sharded._scorer = FusedScorer(ue, ie_shard, item_begin=item_begin)
            return sharded._scorer.topk(None, k, defer=defer)
        sharded._scorer = None
        return engine.topk_exact(ue, ie_shard, None, k, item_begin=item_begin)

    def pending_flag(self, sharded, device) -> torch.Tensor:
        """int64 [1] on the device: 1 when this rank's enqueued fallback leaves work for tier 3."""
        sc = sharded._scorer
        if sc is None or sc._pending is None:
            return torch.zeros(1, dtype=torch.int64, device=device)
        f = sc._pending["flags"]
        return ((f[0] > sc._pending["slots"]) | (f[1] > 0)).to(torch.int64).view(1)

    def finalize_topk(self, sharded) -> bool:
        return sharded._scorer.finalize() if sharded._scorer is not None else False

    def merge(self, ids, scores):
        from . import engine
        return engine.merge_topk(ids, scores)


def profile_stages(model, sharded: Optional[ShardedLightGCN], steps: int = 3) -> Dict[str, float]:
    """Per-stage device times (CUDA events on the launching stream) of one hot-path pass on this rank."""
    from . import engine
    from ._lib import call, ptr, stream

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    g = model.graph
    w = model.embeddings.weight.detach()
    n, d = w.shape
    U = model.num_users
    xs = torch.empty_like(w)
    xo = torch.empty_like(w)
    acc = torch.empty_like(w)
    out: Dict[str, float] = {}
    shard = None
    if sharded is not None and hasattr(sharded.backend, "propagate_sharded"):
        sharded.forward(all_rows=False)                      # builds the rank's UserShard
        shard = sharded.backend._shard
        part = torch.empty(n - U, d, dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        call("hnm_lightgcn_prescale", ptr(w), ptr(g.dis), 0.25, ptr(xs), ptr(acc), n, d, stream())
        heavy = ptr(g.heavy_rows) if g.num_heavy else None

        chunks = getattr(model, "_item_chunks", None) if shard is None else None
        if chunks is not None:
            part = torch.empty(n - U, d, dtype=torch.float32, device=w.device)

        def layer():
            # one layer's kernels on this rank (no communication)
            if chunks is not None:
                engine.layer_call(g, xs, xo, acc, 0.25, 0, U, stream())
                for c, sh in enumerate(chunks.chunks):
                    call("hnm_lightgcn_partial", ptr(sh.seg_begin), ptr(sh.seg_end), ptr(g.col), ptr(g.w), ptr(xs),
                         ptr(part), d, U, n, ptr(sh.heavy_rows) if sh.heavy_rows.numel() else None,
                         int(sh.heavy_rows.numel()), sh.num_huge, g.heavy_threshold, 1 if c else 0, stream())
                call("hnm_lightgcn_finish", ptr(part), ptr(xs), ptr(g.dis), 0.25, ptr(xo), ptr(acc), U, n - U, d, stream())
                return
            if shard is None:
                engine.layer_call(g, xs, xo, acc, 0.25, 0, n, stream())
                return
            if shard.u1 > shard.u0:
                engine.layer_call(g, xs, xo, acc, 0.25, shard.u0, shard.u1, stream())
            call("hnm_lightgcn_partial", ptr(shard.seg_begin), ptr(shard.seg_end), ptr(g.col), ptr(g.w), ptr(xs),
                 ptr(part), d, U, n, ptr(shard.heavy_rows) if shard.heavy_rows.numel() else None,
                 int(shard.heavy_rows.numel()), shard.num_huge, g.heavy_threshold, 0, stream())
            call("hnm_lightgcn_finish", ptr(part), ptr(xs), ptr(g.dis), 0.25, ptr(xo), ptr(acc), U, n - U, d, stream())
        layer()
        a = ev()
        for _ in range(steps):
            layer()
        b = ev()
        torch.cuda.synchronize()
        out["spmm_layer_ms"] = a.elapsed_time(b) / steps
        a = ev()
        for _ in range(steps):
            call("hnm_lightgcn_prescale", ptr(w), ptr(g.dis), 0.25, ptr(xs), ptr(acc), n, d, stream())
        b = ev()
        torch.cuda.synchronize()
        out["prescale_ms"] = a.elapsed_time(b) / steps
    del xs, xo, acc
    # propagate (all layers, exchanges included when sharded)
    fwd = (lambda: sharded.forward(all_rows=sharded.mode != "users")) if sharded is not None else model.forward
    fwd()
    a = ev()
    for _ in range(steps):
        fwd()
    b = ev()
    torch.cuda.synchronize()
    out["propagate_ms"] = a.elapsed_time(b) / steps
    # scoring stages
    rec = sharded.recommend_all if sharded is not None else model.recommend_all
    rec()
    scorer = sharded._scorer if sharded is not None else model._scorer
    fused = {"fused": 0.0, "rescore": 0.0, "pack": 0.0, "fallback": 0.0}
    uncert = 0
    if scorer is not None:
        for _ in range(steps):
            rec()
            scorer = sharded._scorer if sharded is not None else model._scorer
            scorer.profile = True
            scorer.topk(None, model.top_k)
            for k2 in fused:
                fused[k2] += scorer.stage_ms.get(k2, 0.0) / steps
            uncert = dict(scorer.last_stats)
    out.update({"fused_ms": fused["fused"], "rescore_ms": fused["rescore"], "pack_users_ms": fused["pack"],
                "fallback_ms": fused["fallback"], "uncertified_users": uncert.get("uncertified", 0),
                "scorer_stats": uncert})
    return out
