"""Host-side orchestration of the B200 kernels (thin: allocation + call order).

Each function maps onto one call site of the reference's hot path; the kernels
themselves live in csrc/ behind the C ABI (include/hnm_b200.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream

HEAVY_THRESHOLD = 1024          # CSR rows longer than this are summed by a whole CTA
HUGE_ROW = 8192                 # ... and rows longer than this by a cluster of 8 CTAs (HNM_HUGE_ROW)
EXACT_K_MAX = 256               # hnm_topk_exact limit (XCAP - XT in score_exact.cu)


def num_sms(device=None) -> int:
    return torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device()).multi_processor_count


@dataclass
class Graph:
    """Normalised adjacency of LightGCN.set_graph in device CSR form (lightgcn.py:92-112)."""
    num_nodes: int
    nnz: int
    rowptr: torch.Tensor          # int32 [N+1]
    col: torch.Tensor             # int32 [nnz]
    w: Optional[torch.Tensor]     # fp32 [nnz] raw edge weights in CSR order, None when all ones
    dis: torch.Tensor             # fp32 [N]  deg^-1/2
    heavy_rows: torch.Tensor      # int32 [H], the num_huge very long rows first
    heavy_threshold: int = HEAVY_THRESHOLD
    num_huge: int = 0             # rows with more than HUGE_ROW entries (summed by a CTA cluster)
    short_prefix: int = 0         # rows [0, short_prefix) are short on average (the user rows of a bipartite graph):
                                  # they take the staged 8-rows-per-warp kernel (hnm_lightgcn_layer short_rows = 1)

    @property
    def num_heavy(self) -> int:
        return int(self.heavy_rows.numel())


def build_graph(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], num_nodes: int,
                device: torch.device, heavy_threshold: int = HEAVY_THRESHOLD) -> Graph:
    """LightGCN.set_graph (lightgcn.py:81-112): self loops, row degree, deg^-1/2, (row,col)-sorted CSR."""
    _lib.require_device()
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError("edge_index must have shape [2, num_edges]")
    ei = edge_index.to(device=device, dtype=torch.int64).contiguous()
    m = int(ei.size(1))
    ew = None
    if edge_weight is not None:
        ew = edge_weight.detach().to(device=device, dtype=torch.float32).contiguous()
        if ew.numel() != m:
            raise ValueError("edge_weight must have one entry per edge")
    nnz = m + num_nodes
    rowptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=device)
    col = torch.empty(nnz, dtype=torch.int32, device=device)
    w = torch.empty(nnz, dtype=torch.float32, device=device) if ew is not None else None
    dis = torch.empty(num_nodes, dtype=torch.float32, device=device)
    heavy = torch.empty(num_nodes, dtype=torch.int32, device=device)
    ws_bytes = _lib.load().hnm_graph_build_workspace_bytes(num_nodes, m, 1 if ew is not None else 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    n_heavy = C.c_int32(0)
    with torch.cuda.device(device):
        call("hnm_graph_build", ptr(ei[0]), ptr(ei[1]), ptr(ew), m, num_nodes, ptr(rowptr), ptr(col), ptr(w),
             ptr(dis), heavy_threshold, ptr(heavy), C.addressof(n_heavy), ptr(ws), ws_bytes, stream())
    heavy_rows = torch.sort(heavy[: n_heavy.value]).values
    num_huge = 0
    if heavy_rows.numel():
        idx = heavy_rows.long()
        huge = (rowptr[idx + 1] - rowptr[idx]) > HUGE_ROW
        num_huge = int(huge.sum())
        heavy_rows = torch.cat([heavy_rows[huge], heavy_rows[~huge]])
    return Graph(num_nodes, nnz, rowptr, col, w, dis, heavy_rows.contiguous(), heavy_threshold, num_huge)


def layer_call(graph: Graph, cur: torch.Tensor, nxt: Optional[torch.Tensor], acc: torch.Tensor, alpha: float, r0: int,
          r1: int, s) -> None:
    """One hnm_lightgcn_layer call over rows [r0, r1), split at graph.short_prefix so that the short (user) rows
    take the staged kernel and the long (item) rows the row-at-a-time one."""
    n, d = cur.shape
    parts = [(r0, r1, 0)]
    sp = graph.short_prefix
    if r0 < sp:
        parts = [(r0, min(r1, sp), 1)] + ([(sp, r1, 0)] if r1 > sp else [])
    for a, b, short in parts:
        if b > a:
            call("hnm_lightgcn_layer", ptr(graph.rowptr), ptr(graph.col), ptr(graph.w), ptr(graph.dis), ptr(cur),
                 None if nxt is None else ptr(nxt), ptr(acc), float(alpha), n, d, a, b,
                 ptr(graph.heavy_rows) if graph.num_heavy else None, graph.num_heavy, graph.num_huge,
                 graph.heavy_threshold, short, s)


def propagate(graph: Graph, e0: torch.Tensor, alphas: Sequence[float], num_layers: int,
              row_ranges: Optional[Sequence[Tuple[int, int]]] = None, exchange=None,
              exchange_final=None, item_chunks: Optional["ItemChunks"] = None) -> torch.Tensor:
    """LightGCN.forward (lightgcn.py:147-158): returns final [N, d] = sum_l alpha_l * A_hat^l E0.

    row_ranges/exchange serve the row-sharded multi-GPU form: this rank computes the listed
    row ranges of every layer and ``exchange(buf)`` makes all rows of ``buf`` visible
    (an allgather of row slices) before the next layer gathers from it; ``exchange_final`` does the
    same for the returned layer sum (defaults to ``exchange``; a user-sharded scorer only needs the
    item rows of it).
    """
    _lib.require_device()
    if not e0.is_cuda:
        raise RuntimeError("LightGCN parameters must live on a CUDA device; there is no CPU path")
    e0 = e0.detach().to(torch.float32).contiguous()
    n, d = e0.shape
    if n != graph.num_nodes:
        raise ValueError("embedding rows != graph nodes")
    ranges = list(row_ranges) if row_ranges is not None else [(0, n)]
    if item_chunks is not None and row_ranges is None and exchange is None and d % 4 == 0:
        return _propagate_chunked(graph, e0, alphas, num_layers, item_chunks)
    acc = torch.empty_like(e0)
    xs_a = torch.empty_like(e0)
    xs_b = torch.empty_like(e0) if num_layers > 1 else None
    with torch.cuda.device(e0.device):
        s = stream()
        call("hnm_lightgcn_prescale", ptr(e0), ptr(graph.dis), float(alphas[0]), ptr(xs_a), ptr(acc), n, d, s)
        cur, nxt = xs_a, xs_b
        for layer in range(1, num_layers + 1):
            last = layer == num_layers
            for r0, r1 in ranges:
                layer_call(graph, cur, None if last else nxt, acc, alphas[layer], r0, r1, s)
            if not last:
                if exchange is not None:
                    exchange(nxt)
                cur, nxt = nxt, cur
        final_x = exchange_final if exchange_final is not None else exchange
        if final_x is not None:
            final_x(acc)
    return acc


def exclusion_csr(user_ids: torch.Tensor, filter_items: Optional[Dict[int, set]], device) -> Tuple[
        Optional[torch.Tensor], Optional[torch.Tensor]]:
    """filter_items dict (lightgcn.py:349-353) -> CSR over the listed users, item ids sorted ascending."""
    if filter_items is None:
        return None, None
    ptrs: List[int] = [0]
    items: List[int] = []
    for uid in user_ids.tolist():
        if uid in filter_items:
            items.extend(sorted(int(i) for i in filter_items[uid]))
        ptrs.append(len(items))
    if not items:
        return None, None
    return (torch.tensor(ptrs, dtype=torch.int64, device=device),
            torch.tensor(items, dtype=torch.int64, device=device))


def slice_csr(ptr_all: torch.Tensor, items_all: torch.Tensor, user_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rows `user_ids` of a CSR over all users -> a CSR over the listed users (device ops only)."""
    lo, hi = ptr_all[user_ids], ptr_all[user_ids + 1]
    lens = hi - lo
    out_ptr = torch.zeros(user_ids.numel() + 1, dtype=torch.int64, device=ptr_all.device)
    torch.cumsum(lens, 0, out=out_ptr[1:])
    total = int(out_ptr[-1])
    if total == 0:
        return out_ptr, torch.zeros(1, dtype=torch.int64, device=ptr_all.device)
    row = torch.repeat_interleave(torch.arange(user_ids.numel(), device=ptr_all.device), lens)
    pos = torch.arange(total, device=ptr_all.device) - out_ptr[row] + lo[row]
    return out_ptr, items_all[pos].contiguous()


def history_csr(graph: Graph, num_users: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """The purchased-item filter of the serving path (scripts/serve.py:174-177,350-352) straight from the graph:
    user u's row of the adjacency holds its self loop and the item nodes it bought, so the exclusion lists are
    the user block of the CSR minus the self loops, item NODE ids turned into item indices.  Sorted per user
    (the CSR is), repeat purchases appear more than once (harmless).  Bipartite graphs only."""
    if not is_bipartite(graph, num_users):
        raise ValueError("history_csr needs a bipartite user-item graph")
    split = int(graph.rowptr[num_users])
    col = graph.col[:split]
    items = (col[col >= num_users] - num_users).to(torch.int64).contiguous()
    ptr_ = graph.rowptr[: num_users + 1].to(torch.int64) - torch.arange(num_users + 1, device=col.device)
    return ptr_.contiguous(), items


def _norm_ids(ids: Optional[torch.Tensor], limit: int, device) -> Optional[torch.Tensor]:
    if ids is None:
        return None
    ids = ids.to(device=device, dtype=torch.int64).contiguous().view(-1)
    if ids.numel():
        lo, hi = int(ids.min()), int(ids.max())
        if lo < -limit or hi >= limit:
            raise IndexError(f"index out of range in self (valid: [{-limit}, {limit - 1}])")
        if lo < 0:
            ids = torch.where(ids < 0, ids + limit, ids)
    return ids


def pair_scores(user_emb, item_emb, user_ids, item_ids) -> torch.Tensor:
    """LightGCN.predict tail (lightgcn.py:180-184).  Indices are validated here (as for every other entry point),
    so the kernel call itself stays asynchronous."""
    _lib.require_device()
    dev = user_emb.device
    u = _norm_ids(user_ids, user_emb.size(0), dev)
    i = _norm_ids(item_ids, item_emb.size(0), dev)
    if u.numel() != i.numel():
        raise RuntimeError("user_ids and item_ids must have the same length")
    out = torch.empty(u.numel(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("hnm_pair_scores", ptr(user_emb), ptr(item_emb), ptr(u), ptr(i), u.numel(), user_emb.size(1),
             user_emb.size(0), item_emb.size(0), ptr(out), None, stream())
    return out


def score_all_items(user_emb, item_emb, user_ids) -> torch.Tensor:
    """LightGCN.predict_all_items tail (lightgcn.py:199-202): [B, I] fp32."""
    _lib.require_device()
    dev = user_emb.device
    u = _norm_ids(user_ids, user_emb.size(0), dev)
    b = u.numel()
    out = torch.empty(b, item_emb.size(0), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("hnm_score_all_items", ptr(user_emb), ptr(item_emb), ptr(u), b, item_emb.size(0), user_emb.size(1),
             ptr(out), stream())
    return out


def topk_exact(user_emb, item_emb, user_ids: Optional[torch.Tensor], k: int,
               excl: Tuple[Optional[torch.Tensor], Optional[torch.Tensor]] = (None, None),
               item_begin: int = 0, batch: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact canonical top-k (fp64 scores; score desc, id asc).  item_emb holds the local shard
    [item_begin, item_begin + rows)."""
    _lib.require_device()
    dev = user_emb.device
    u = _norm_ids(user_ids, user_emb.size(0), dev)
    b = u.numel() if u is not None else (batch if batch is not None else user_emb.size(0))
    # few users (the fallback of the tensor-core path): split the item range so the launch fills the GPU
    n_items = int(item_emb.size(0))
    ctas = max(1, (b + 7) // 8)
    splits = max(1, min(16, (2 * num_sms(dev) + ctas - 1) // ctas, n_items // max(1024, k)))
    ids = torch.empty(splits, b, k, dtype=torch.int64, device=dev)
    sc = torch.empty(splits, b, k, dtype=torch.float64, device=dev)
    if b == 0:
        return ids[0], sc[0]
    with torch.cuda.device(dev):
        call("hnm_topk_exact", ptr(user_emb), ptr(item_emb), ptr(u), b, item_begin, item_begin + n_items,
             user_emb.size(1), k, ptr(excl[0]), ptr(excl[1]), splits, ptr(ids), ptr(sc), stream())
    if splits == 1:
        return ids[0], sc[0]
    return merge_topk(ids, sc)


def merge_topk(ids: torch.Tensor, scores: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[G, B, k] per-shard sorted lists -> global [B, k]."""
    _lib.require_device()
    g, b, k = ids.shape
    out_i = torch.empty(b, k, dtype=torch.int64, device=ids.device)
    out_s = torch.empty(b, k, dtype=torch.float64, device=ids.device)
    with torch.cuda.device(ids.device):
        call("hnm_merge_topk", ptr(ids.contiguous()), ptr(scores.contiguous()), g, b, k, ptr(out_i), ptr(out_s),
             stream())
    return out_i, out_s


# ----------------------------------------------------------------------------- user-sharded propagation
@dataclass
class UserShard:
    """What one rank needs to propagate with the USERS partitioned: its user range and, for every item
    row, the sub-range of CSR entries whose column is one of its users (contiguous: columns are sorted)."""
    u0: int
    u1: int
    seg_begin: torch.Tensor       # int32 [I]
    seg_end: torch.Tensor         # int32 [I]
    heavy_rows: torch.Tensor      # int32, item NODE ids whose sub-range is long (very long ones first)
    num_huge: int


def _lower_bound(col: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, target: int) -> torch.Tensor:
    """Per row, first position p in [lo, hi) with col[p] >= target (vectorised bisection)."""
    lo, hi = lo.clone(), hi.clone()
    n = col.numel()
    for _ in range(34):
        active = lo < hi
        if not bool(active.any()):
            break
        mid = (lo + hi) // 2
        less = col[mid.clamp(max=n - 1)] < target
        lo = torch.where(active & less, mid + 1, lo)
        hi = torch.where(active & ~less, mid, hi)
    return lo


def make_user_shard(graph: Graph, num_users: int, num_items: int, u0: int, u1: int) -> UserShard:
    rp = graph.rowptr.long()
    lo, hi = rp[num_users:num_users + num_items], rp[num_users + 1:num_users + num_items + 1]
    b = _lower_bound(graph.col, lo, hi, u0)
    e = _lower_bound(graph.col, lo, hi, u1)
    length = e - b
    rows = torch.arange(num_users, num_users + num_items, device=rp.device)
    heavy = length > graph.heavy_threshold
    huge = length > HUGE_ROW
    heavy_rows = torch.cat([rows[huge], rows[heavy & ~huge]]).to(torch.int32).contiguous()
    return UserShard(u0, u1, b.to(torch.int32).contiguous(), e.to(torch.int32).contiguous(), heavy_rows,
                     int(huge.sum()))


def is_bipartite(graph: Graph, num_users: int) -> bool:
    """True when user rows only point at items and item rows only at users (self loops aside) -- what the
    user-partitioned and the chunked propagation assume.  One device reduction, cached on the graph."""
    cache = graph.__dict__.setdefault("_bipartite", {})
    if num_users not in cache:
        n = graph.num_nodes
        split = int(graph.rowptr[num_users])
        in_user_rows = int((graph.col[:split] >= num_users).sum())      # item columns (+ nothing else allowed)
        in_item_rows = int((graph.col[split:] >= num_users).sum())      # must be exactly the self loops
        cache[num_users] = (in_user_rows == split - num_users) and (in_item_rows == n - num_users)
    return cache[num_users]


@dataclass
class ItemChunks:
    """Sub-ranges of the item rows per L2-sized chunk of users (see hnm_lightgcn_partial in the header)."""
    num_users: int
    chunks: List[UserShard]


def make_item_chunks(graph: Graph, num_users: int, num_items: int, dim: int,
                     num_chunks: Optional[int] = None) -> Optional[ItemChunks]:
    """None unless chunking was asked for (num_chunks > 1 or HNM_SPMM_CHUNKS) and the graph is bipartite."""
    import os
    if num_chunks is None and os.environ.get("HNM_SPMM_CHUNKS"):
        num_chunks = int(os.environ["HNM_SPMM_CHUNKS"])
    if num_chunks is None:
        # measured at the H&M shape (profiles/r1_spmm_notes.md): 1.76 ms per layer in one pass, 1.93 / 1.95 /
        # 2.64 ms with 4 / 8 / 16 chunks -- the pass is bound by the L2 -> SM gather traffic, not by the HBM
        # re-reads the chunks remove, so the chunked form stays opt-in (HNM_SPMM_CHUNKS, or num_chunks here)
        num_chunks = 1
    if num_chunks <= 1 or num_users + num_items != graph.num_nodes or not is_bipartite(graph, num_users):
        return None
    base, rem = divmod(num_users, num_chunks)
    out, a = [], 0
    for c in range(num_chunks):
        b = a + base + (1 if c < rem else 0)
        out.append(make_user_shard(graph, num_users, num_items, a, b))
        a = b
    return ItemChunks(num_users, out)


def _propagate_chunked(graph: Graph, e0: torch.Tensor, alphas: Sequence[float], num_layers: int,
                       ic: ItemChunks) -> torch.Tensor:
    """LightGCN.forward on one GPU with the item rows summed chunk of users by chunk of users, so that the
    user rows a chunk gathers stay in the L2 (each leaves HBM once per layer instead of ~6 times)."""
    n, d = e0.shape
    U = ic.num_users
    I = n - U
    acc = torch.empty_like(e0)
    xs_a = torch.empty_like(e0)
    xs_b = torch.empty_like(e0) if num_layers > 1 else None
    part = torch.empty(I, d, dtype=torch.float32, device=e0.device)
    with torch.cuda.device(e0.device):
        s = stream()
        call("hnm_lightgcn_prescale", ptr(e0), ptr(graph.dis), float(alphas[0]), ptr(xs_a), ptr(acc), n, d, s)
        cur, nxt = xs_a, xs_b
        heavy = ptr(graph.heavy_rows) if graph.num_heavy else None
        for layer in range(1, num_layers + 1):
            last = layer == num_layers
            layer_call(graph, cur, None if last else nxt, acc, alphas[layer], 0, U, s)
            for c, sh in enumerate(ic.chunks):
                call("hnm_lightgcn_partial", ptr(sh.seg_begin), ptr(sh.seg_end), ptr(graph.col), ptr(graph.w), ptr(cur),
                     ptr(part), d, U, n, ptr(sh.heavy_rows) if sh.heavy_rows.numel() else None,
                     int(sh.heavy_rows.numel()), sh.num_huge, graph.heavy_threshold, 1 if c else 0, s)
            call("hnm_lightgcn_finish", ptr(part), ptr(cur), ptr(graph.dis), float(alphas[layer]),
                 None if last else ptr(nxt), ptr(acc), U, I, d, s)
            if not last:
                cur, nxt = nxt, cur
    return acc


def propagate_user_sharded(graph: Graph, shard: UserShard, e0: torch.Tensor, alphas: Sequence[float],
                           num_layers: int, num_users: int, allreduce_items) -> torch.Tensor:
    """LightGCN.forward with the users partitioned over ranks (one process per GPU).

    Every rank keeps all item rows and its own user rows.  Per layer: its user rows gather from the
    item block as usual; for the item rows it sums only over its own users (hnm_lightgcn_partial),
    ``allreduce_items(partial)`` adds the ranks' partial sums in place (27 MB at the H&M shape, instead
    of all-gathering the 351 MB user block), and hnm_lightgcn_finish normalises.  Returns acc [N, d]:
    valid for the item rows and for user rows [u0, u1).
    """
    _lib.require_device()
    e0 = e0.detach().to(torch.float32).contiguous()
    n, d = e0.shape
    U = num_users
    I = n - U
    if d % 4:
        raise ValueError("user-sharded propagation needs embedding_dim % 4 == 0")
    dev = e0.device
    u0, u1 = shard.u0, shard.u1
    acc = torch.empty_like(e0)
    xs_a = torch.empty_like(e0)
    xs_b = torch.empty_like(e0) if num_layers > 1 else None
    part = torch.empty(I, d, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        s = stream()
        if u1 > u0:
            call("hnm_lightgcn_prescale", ptr(e0[u0:u1]), ptr(graph.dis[u0:u1]), float(alphas[0]), ptr(xs_a[u0:u1]),
                 ptr(acc[u0:u1]), u1 - u0, d, s)
        call("hnm_lightgcn_prescale", ptr(e0[U:]), ptr(graph.dis[U:]), float(alphas[0]), ptr(xs_a[U:]), ptr(acc[U:]),
             I, d, s)
        cur, nxt = xs_a, xs_b
        heavy = ptr(graph.heavy_rows) if graph.num_heavy else None
        sh_heavy = ptr(shard.heavy_rows) if shard.heavy_rows.numel() else None
        for layer in range(1, num_layers + 1):
            last = layer == num_layers
            call("hnm_lightgcn_partial", ptr(shard.seg_begin), ptr(shard.seg_end), ptr(graph.col), ptr(graph.w),
                 ptr(cur), ptr(part), d, U, n, sh_heavy, int(shard.heavy_rows.numel()), shard.num_huge,
                 graph.heavy_threshold, 0, s)
            # the all-reduce may be asynchronous (it then returns a handle): this rank's user rows, which
            # need neither `part` nor the other ranks, are gathered while the partial sums travel
            pending = allreduce_items(part)
            if u1 > u0:
                layer_call(graph, cur, None if last else nxt, acc, alphas[layer], u0, u1, s)
            if pending is not None:
                pending.wait()
            call("hnm_lightgcn_finish", ptr(part), ptr(cur), ptr(graph.dis), float(alphas[layer]),
                 None if last else ptr(nxt), ptr(acc), U, I, d, s)
            if not last:
                cur, nxt = nxt, cur
    return acc


class PeerBuffers:
    """Symmetric (peer-mapped) buffers of the NVLink exchange: the two pre-scaled tables, the layer sum and the
    staging buffer for the partial sums, allocated once per (graph, world) with torch's symmetric memory
    (cuMem allocations whose handles the ranks exchange; every rank sees every rank's buffer).  Item row i is
    owned by rank i // rows_per_owner."""

    def __init__(self, num_nodes: int, num_items: int, dim: int, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rpo = -(-num_items // self.world)
        dev = torch.device("cuda", torch.cuda.current_device())
        pg = group if group is not None else dist.group.WORLD

        def make(*shape):
            t = symm.empty(*shape, dtype=torch.float32, device=dev)
            return t, symm.rendezvous(t, pg)

        self.xs_a, self.h_a = make(num_nodes, dim)
        self.xs_b, self.h_b = make(num_nodes, dim)
        self.acc, self.h_acc = make(num_nodes, dim)
        self.stage, self.h_stage = make(self.world * self.rpo, dim)

    def handle(self, t: torch.Tensor):
        return {id(self.xs_a): self.h_a, id(self.xs_b): self.h_b, id(self.acc): self.h_acc}[id(t)]


def propagate_user_sharded_peer(graph: Graph, shard: UserShard, e0: torch.Tensor, alphas: Sequence[float],
                                num_layers: int, num_users: int, pb: PeerBuffers) -> torch.Tensor:
    """propagate_user_sharded with the exchange inside the kernels: the partial-sum kernel stores every item
    row's partial straight into its owner's staging buffer over NVLink, the owner reduces in rank order, finishes
    the row and writes it into every rank's table (hnm_lightgcn_partial_peer / hnm_lightgcn_finish_peer), with
    one cross-rank barrier after each of the two.  No NCCL call and no replicated finish pass on the data path.
    Returns the symmetric layer-sum buffer (reused by the next call): valid for all item rows and for user rows
    [u0, u1)."""
    _lib.require_device()
    e0 = e0.detach().to(torch.float32).contiguous()
    n, d = e0.shape
    U = num_users
    I = n - U
    u0, u1 = shard.u0, shard.u1
    acc, xs_a, xs_b = pb.acc, pb.xs_a, pb.xs_b
    with torch.cuda.device(e0.device):
        s = stream()
        # No barrier is needed here: a peer writes into this rank's buffers only after a barrier of the NEW step that
        # this rank has reached, i.e. after everything it enqueued for the previous step (scoring included) is done.
        if u1 > u0:
            call("hnm_lightgcn_prescale", ptr(e0[u0:u1]), ptr(graph.dis[u0:u1]), float(alphas[0]), ptr(xs_a[u0:u1]),
                 ptr(acc[u0:u1]), u1 - u0, d, s)
        call("hnm_lightgcn_prescale", ptr(e0[U:]), ptr(graph.dis[U:]), float(alphas[0]), ptr(xs_a[U:]), ptr(acc[U:]),
             I, d, s)
        cur, nxt = xs_a, xs_b
        heavy = ptr(graph.heavy_rows) if graph.num_heavy else None
        sh_heavy = ptr(shard.heavy_rows) if shard.heavy_rows.numel() else None
        for layer in range(1, num_layers + 1):
            last = layer == num_layers
            call("hnm_lightgcn_partial_peer", ptr(shard.seg_begin), ptr(shard.seg_end), ptr(graph.col), ptr(graph.w),
                 ptr(cur), pb.h_stage.buffer_ptrs_dev, pb.rpo, pb.rank, d, U, n, sh_heavy,
                 int(shard.heavy_rows.numel()), shard.num_huge, graph.heavy_threshold, s)
            if u1 > u0:       # this rank's user rows need neither the partial sums nor the other ranks
                layer_call(graph, cur, None if last else nxt, acc, alphas[layer], u0, u1, s)
            pb.h_stage.barrier(channel=0)                 # every rank's partial sums have landed
            call("hnm_lightgcn_finish_peer", ptr(pb.stage), pb.world, pb.rpo, pb.rank, ptr(cur), ptr(graph.dis),
                 float(alphas[layer]), None if last else pb.handle(nxt).buffer_ptrs_dev, ptr(acc),
                 pb.h_acc.buffer_ptrs_dev if last else None, U, I, d, s)
            pb.h_stage.barrier(channel=1)                 # ... and every owner's finished rows
            if not last:
                cur, nxt = nxt, cur
    return acc
