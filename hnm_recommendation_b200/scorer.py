"""Full-catalog scorer on the tensor cores: pack -> fused score/select -> exact rescore -> fallback.

Replaces ``torch.matmul(user_embeds, item_embeddings.t())`` + ``torch.topk`` of
src/models/lightgcn.py:202,356.  Exactness does not rest on the fp16 tensor-core
scores: they only nominate candidates; every returned list is ranked by exact
fp64 scores and carries a certificate, and users that cannot be certified are
recomputed by the exact SIMT kernel (hnm_topk_exact).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib, engine
from ._lib import call, ptr, stream

USER_BLOCK = 128
ITEM_TILE = 128
CAND_CAP = 192                 # candidate entries (32-column chunks) per user; ~100 on average at the H&M shape
CAND_CAP_BIG = 384             # ... the count grows like (k + margin) ln(#item tiles): catalogs beyond ~250 k items
CAND_CAP_MAX = 1024            # HNM_FUSED_CAND_MAX
CAND_WORDS = 5                 # HNM_FUSED_CAND_BYTES / 4 (16 bytes of group maxima + 4 bytes of column per entry)
SIG_WORDS = 32                 # HNM_FUSED_SIG_WORDS
K_MAX = 16
FUSED_DIMS = (64, 128, 256)    # embedding dimensions of the tensor-core path (K chunks of 64)
SEL_MARGIN = 3                 # tau tracks the (k + margin)-th best bucket maximum
SEL_MARGIN_BIG = 5             # ... catalogs beyond ~250 k items: the top scores lie closer together (profiles/r2_margin_sweep_configs4.jsonl)
MAX_USERS_PER_LAUNCH = 1 << 21
TIER2_MIN_USERS = 32            # fewer uncertified users go straight to the exact kernel
TIER2_SLOTS = 512               # smallest second pass when it is enqueued without knowing the count (no host sync)
TIER2_MARGIN = 12               # tau of the second pass tracks the (k + 12)-th best bucket maximum


class FusedScorer:
    """Holds the packed item shard; scores any list of users against it."""

    @staticmethod
    def supports(dim: int, k: int, num_items: int) -> bool:
        return dim in FUSED_DIMS and 1 <= k <= K_MAX and num_items >= 2 * ITEM_TILE

    def __init__(self, user_emb: torch.Tensor, item_emb: torch.Tensor, item_begin: int = 0,
                 center: bool = True, sel_margin: Optional[int] = None, tier2_margin: int = TIER2_MARGIN,
                 cand_cap: Optional[int] = None):
        _lib.require_device()
        self.user_emb = user_emb.contiguous()
        self.item_emb = item_emb.contiguous()
        self.item_begin = item_begin
        self.tier2_margin = tier2_margin
        dev = self.item_emb.device
        self.num_items = int(self.item_emb.size(0))
        self.dim = int(self.item_emb.size(1))
        if self.dim not in FUSED_DIMS or int(self.user_emb.size(1)) != self.dim:
            raise ValueError(f"the fused scorer takes embedding dimensions {FUSED_DIMS}")
        self.items_padded = (self.num_items + ITEM_TILE - 1) // ITEM_TILE * ITEM_TILE
        big = self.num_items > 250_000
        self.sel_margin = int(sel_margin) if sel_margin is not None else (SEL_MARGIN_BIG if big else SEL_MARGIN)
        self.cand_cap = int(cand_cap) if cand_cap else (CAND_CAP_BIG if big else CAND_CAP)
        if self.cand_cap % 2 or not 2 <= self.cand_cap <= CAND_CAP_MAX:
            raise ValueError(f"cand_cap must be even and at most {CAND_CAP_MAX}")
        # share of the users the sync-free second pass has slots for: near-ties between the k-th score and tau
        # become more frequent with the catalog size and the embedding dimension (0.07 % of the users at the
        # H&M shape, 0.7 % at 1 M items x 256)
        self.tier2_share = 1.0 / 64 if big else 1.0 / 256
        with torch.cuda.device(dev):
            # any fp32 vector is a valid centre; the mean row is the one that shrinks the items most
            self.center = None
            if center:
                self.center = torch.empty(self.dim, dtype=torch.float32, device=dev)
                ws_bytes = int(_lib.load().hnm_column_mean_workspace_bytes(self.dim))
                ws = torch.empty(ws_bytes // 8, dtype=torch.float64, device=dev)
                call("hnm_column_mean", ptr(self.item_emb), self.num_items, self.dim, ptr(self.center), ptr(ws),
                     ws_bytes, stream())
            # {absmax, scale, max ||x - c||^2, -} of the shard: produced and consumed on the device, so the
            # set-up needs no host synchronisation (two .tolist()/.max() round trips in round 1)
            self.item_params = torch.zeros(4, dtype=torch.float32, device=dev)
            call("hnm_absmax", ptr(self.item_emb), self.item_emb.numel(), ptr(self.center), self.dim,
                 self.item_params.data_ptr(), stream())
            self.items_f16 = torch.empty(self.items_padded, self.dim, dtype=torch.float16, device=dev)
            call("hnm_score_pack_items", ptr(self.item_emb), self.num_items, self.items_padded, self.dim,
                 ptr(self.center), ptr(self.item_params), ptr(self.items_f16), stream())
        self.last_stats: Dict[str, int] = {}
        self.profile = False                       # record CUDA events around each stage of topk()
        self.stage_ms: Dict[str, float] = {}
        self._events = []

    def topk(self, user_ids: Optional[torch.Tensor], k: int, filter_items=None,
             fallback: bool = True, out_host: Optional[torch.Tensor] = None,
             chunk_users: Optional[int] = None, defer: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """ids [B,k] int64 (global item indices), scores [B,k] fp64; canonical (score desc, id asc).

        Three tiers, each only for the users the previous one could not certify:
        1. tensor-core nomination with tau tracking the (k + sel_margin)-th best score, exact rescoring;
        2. the same with a wider margin (k + 12): near-ties between the k-th score and tau disappear;
        3. exact fp64 brute force (hnm_topk_exact): exact ties, overflowing lists, anything else.

        filter_items: the reference's ``{user_id: set(item ids)}`` dict (lightgcn.py:349-353), or a device CSR
        ``(ptr int64 [U_all + 1], items int64 sorted per user)`` over ALL users of the table (engine.history_csr:
        the purchased-item filter at full scale without a Python loop).  Filtered users stay on the tensor
        path: their excluded chunks are kept out of the nomination threshold (hnm_exclusion_signature).

        out_host (pinned int64 [B, k]): the ids are also delivered to the host, chunk of users by chunk of
        users on a copy stream while the next chunk is being scored; the few rows the fallback tiers rewrite
        are patched afterwards.  The call returns with out_host complete.

        Without a filter and without out_host the whole call is enqueued WITHOUT a host synchronisation: the
        second tier runs on a fixed number of slots filled by a device-side compaction of the uncertified users,
        and the two counters that say whether anything is left (more uncertified users than slots, or users the
        second tier could not certify either) are read once, at the end -- by this call, or, with defer=True, by
        ``finalize()``, which the caller invokes after it has enqueued whatever follows (the sharded form
        launches its result all-gather first).  ``finalize()`` returns True when it had to patch rows.
        """
        dev = self.item_emb.device
        uids = engine._norm_ids(user_ids, self.user_emb.size(0), dev)
        total = uids.numel() if uids is not None else int(self.user_emb.size(0))
        ids_full = torch.empty(total + 1, k, dtype=torch.int64, device=dev)      # + one spare row, see _enqueue_fallback
        sc_full = torch.empty(total + 1, k, dtype=torch.float64, device=dev)
        ids, sc = ids_full[:total], sc_full[:total]
        cert = torch.empty(total, dtype=torch.int32, device=dev)
        excl = (None, None)
        full_csr = isinstance(filter_items, tuple)
        if filter_items is not None:
            if uids is None:
                uids = torch.arange(total, device=dev)
            excl = self._exclusions(uids, filter_items)
        self._events = []
        sel = min(32, k + self.sel_margin)
        step = MAX_USERS_PER_LAUNCH
        copy_stream = None
        if out_host is not None:
            if tuple(out_host.shape) != (total, k) or out_host.dtype != torch.int64 or not out_host.is_pinned():
                raise ValueError("out_host must be a pinned int64 tensor of shape [users, k]")
            # four chunks: the device-to-host copy of one hides behind the scoring of the next
            step = chunk_users or max(USER_BLOCK * 256, -(-total // 4 // USER_BLOCK) * USER_BLOCK)
            copy_stream = self._copy_stream()
        for b0 in range(0, total, step):
            b1 = min(total, b0 + step)
            self._launch(uids, b0, b1, k, sel, excl, ids, sc, cert)
            if copy_stream is not None:
                copy_stream.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(copy_stream):
                    out_host[b0:b1].copy_(ids[b0:b1], non_blocking=True)
        self.last_stats = {"users": total, "uncertified": 0, "tier2": 0, "tier3": 0}
        why = None
        self._pending = None
        self._mark("fallback_begin")
        if fallback and total and filter_items is None and out_host is None and not self.profile:
            self._enqueue_fallback(uids, total, k, ids_full, sc_full, cert)
            self._mark("fallback_end")
            if not defer:
                self.finalize()
            return ids, sc
        if fallback and total:
            bad = (cert != 1).nonzero().view(-1)
            n_bad = int(bad.numel())
            self.last_stats["uncertified"] = n_bad
            why = cert[bad] if (n_bad and self.profile) else None
            if n_bad:
                bad_uids = bad if uids is None else uids[bad]
                sub_excl = self._exclusions(bad_uids, filter_items) if excl[0] is not None else (None, None)
                sel2 = min(32, k + self.tier2_margin)
                # the second tensor-core pass slices the item range of its few user tiles over the CTAs
                # (SplitPlan in csrc/score_fused.cu), so it is cheap for any count; only a handful of users
                # go straight to the exact kernel (which splits the catalog over CTAs too)
                if sel2 > sel and n_bad >= TIER2_MIN_USERS:
                    ids2 = torch.empty(n_bad, k, dtype=torch.int64, device=dev)
                    sc2 = torch.empty(n_bad, k, dtype=torch.float64, device=dev)
                    cert2 = torch.empty(n_bad, dtype=torch.int32, device=dev)
                    self._launch(bad_uids, 0, n_bad, k, sel2, sub_excl, ids2, sc2, cert2, mark=False)
                    # every tier-2 row is written back (the still uncertified ones are overwritten by tier 3):
                    # two scatters and one host read instead of a mask, four gathers and two scatters
                    ids.index_copy_(0, bad, ids2)
                    sc.index_copy_(0, bad, sc2)
                    self.last_stats["tier2"] = n_bad
                    rest = (cert2 != 1).nonzero().view(-1)
                    bad = bad[rest]
                    bad_uids = bad_uids[rest]
                    if bad.numel() and excl[0] is not None:
                        sub_excl = self._exclusions(bad_uids, filter_items)
                if bad.numel():
                    self.last_stats["tier3"] = int(bad.numel())
                    e_ids, e_sc = engine.topk_exact(self.user_emb, self.item_emb, bad_uids, k, sub_excl,
                                                    item_begin=self.item_begin)
                    ids[bad] = e_ids
                    sc[bad] = e_sc
        self._mark("fallback_end")
        if copy_stream is not None:
            fixed = (cert != 1).nonzero().view(-1) if (fallback and total) else None
            if fixed is not None and fixed.numel():
                rows = ids[fixed].cpu()                        # after the fallback tiers, on the compute stream
                copy_stream.synchronize()
                out_host[fixed.cpu()] = rows
            else:
                copy_stream.synchronize()
        if self.profile:
            torch.cuda.synchronize(dev)
            if fallback and total and why is not None:
                self.last_stats.update({"overflow": int((why & 2).ne(0).sum()), "many_groups": int((why & 4).ne(0).sum()),
                                        "few_contenders": int((why & 8).ne(0).sum()),
                                        "many_contenders": int((why & 16).ne(0).sum())})
            ms: Dict[str, float] = {}
            for (n0, e0), (n1, e1) in zip(self._events[:-1], self._events[1:]):
                if n0.endswith("_begin") and n1 == n0[:-6] + "_end":
                    ms[n0[:-6]] = ms.get(n0[:-6], 0.0) + e0.elapsed_time(e1)
            self.stage_ms = ms
        return ids, sc

    def _enqueue_fallback(self, uids, total, k, ids_full, sc_full, cert) -> None:
        """Tier 2 on a fixed number of slots, no host round trip (see topk).  ids_full / sc_full carry one spare
        row at index `total` that absorbs the write-back of the unused slots."""
        dev = self.item_emb.device
        slots = max(TIER2_SLOTS, int(total * self.tier2_share))
        slots = min(total, -(-slots // (2 * USER_BLOCK)) * 2 * USER_BLOCK)          # whole pairs of user tiles
        sel2 = min(32, k + self.tier2_margin)
        bad = torch.nonzero_static(cert != 1, size=slots, fill_value=-1).view(-1)       # device-side compaction
        valid = bad >= 0
        safe = bad.clamp(min=0)
        bad_uids = safe if uids is None else uids[safe]
        ids2 = torch.empty(slots, k, dtype=torch.int64, device=dev)
        sc2 = torch.empty(slots, k, dtype=torch.float64, device=dev)
        cert2 = torch.empty(slots, dtype=torch.int32, device=dev)
        self._launch(bad_uids, 0, slots, k, sel2, (None, None), ids2, sc2, cert2, mark=False)
        tgt = torch.where(valid, bad, torch.full_like(bad, total))
        ids_full.index_copy_(0, tgt, ids2)
        sc_full.index_copy_(0, tgt, sc2)
        flags = torch.stack([(cert != 1).sum(), ((cert2 != 1) & valid).sum()])
        self._pending = dict(uids=uids, total=total, k=k, ids=ids_full[:total], sc=sc_full[:total], cert=cert, bad=bad,
                             valid=valid, cert2=cert2, flags=flags, slots=slots)

    def finalize(self) -> bool:
        """Read the two counters of the enqueued fallback (the call's only host synchronisation) and handle what
        the fixed-size second tier left: users beyond its slots, users it could not certify (tier 3)."""
        p, self._pending = self._pending, None
        if p is None:
            return False
        n_bad, n_left = (int(x) for x in p["flags"].tolist())
        self.last_stats.update({"uncertified": n_bad, "tier2": min(n_bad, p["slots"]), "tier3": 0})
        if n_bad <= p["slots"] and n_left == 0:
            return False
        dev = self.item_emb.device
        ids, sc, cert, uids, k = p["ids"], p["sc"], p["cert"], p["uids"], p["k"]
        fixed = torch.zeros(p["total"], dtype=torch.bool, device=dev)
        ok2 = p["valid"] & (p["cert2"] == 1)
        fixed[p["bad"][ok2]] = True
        rest = ((cert != 1) & ~fixed).nonzero().view(-1)
        if rest.numel():
            rest_uids = rest if uids is None else uids[rest]
            e_ids, e_sc = engine.topk_exact(self.user_emb, self.item_emb, rest_uids, k, (None, None),
                                            item_begin=self.item_begin)
            ids[rest] = e_ids
            sc[rest] = e_sc
            self.last_stats["tier3"] = int(rest.numel())
        return True

    def _exclusions(self, uids: torch.Tensor, filter_items):
        """Exclusion CSR over the listed users: from the reference's dict, or sliced out of an all-users CSR."""
        dev = self.item_emb.device
        if isinstance(filter_items, tuple):
            return engine.slice_csr(filter_items[0], filter_items[1], uids)
        return engine.exclusion_csr(uids, filter_items, dev)

    def _copy_stream(self) -> torch.cuda.Stream:
        if getattr(self, "_cstream", None) is None:
            self._cstream = torch.cuda.Stream(device=self.item_emb.device)
        return self._cstream

    def _mark(self, name: str) -> None:
        if self.profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._events.append((name, ev))

    def _launch(self, uids, b0, b1, k, sel, excl, ids, sc, cert, mark: bool = True) -> None:
        dev = self.item_emb.device
        n = b1 - b0
        padded = (n + USER_BLOCK - 1) // USER_BLOCK * USER_BLOCK
        note = self._mark if mark else (lambda name: None)
        with torch.cuda.device(dev):
            s = stream()
            users_f16 = torch.empty(padded, self.dim, dtype=torch.float16, device=dev)
            inv_scale = torch.empty(n, dtype=torch.float32, device=dev)      # 1 / (the row's own power of two)
            note("pack_begin")
            if uids is None:
                src = self.user_emb[b0:b1]
                call("hnm_score_pack_users", ptr(src), None, n, padded, self.dim, ptr(users_f16), ptr(inv_scale), s)
                rid = None
                user_base = src
            else:
                rid = uids[b0:b1]
                call("hnm_score_pack_users", ptr(self.user_emb), ptr(rid), n, padded, self.dim, ptr(users_f16),
                     ptr(inv_scale), s)
                user_base = self.user_emb
            cand = torch.empty(n * self.cand_cap * CAND_WORDS, dtype=torch.int32, device=dev)
            count = torch.empty(n, 2, dtype=torch.int32, device=dev)     # one list per thread of the row
            thresh = torch.empty(n, dtype=torch.float32, device=dev)
            ws_bytes = int(_lib.load().hnm_score_topk_fused_workspace_bytes(padded, self.items_padded))
            if ws_bytes < 0:
                raise ValueError("bad padded sizes for the fused scorer")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            note("pack_end")
            ex_ptr = excl[0][b0:b1 + 1] if excl[0] is not None else None
            sig = None
            if ex_ptr is not None:
                sig = torch.empty(n, SIG_WORDS, dtype=torch.int32, device=dev)
                call("hnm_exclusion_signature", ptr(ex_ptr), ptr(excl[1]), n, self.item_begin, self.num_items,
                     ptr(sig), s)
            note("fused_begin")
            call("hnm_score_topk_fused", ptr(users_f16), n, padded, ptr(self.items_f16), self.num_items,
                 self.items_padded, self.dim, sel, ptr(cand), self.cand_cap, ptr(count), ptr(thresh), ptr(sig), ptr(ws),
                 ws_bytes, s)
            note("fused_end")
            note("rescore_begin")
            call("hnm_rescore_topk", ptr(user_base), ptr(self.item_emb), ptr(rid), n, self.dim, self.item_begin,
                 self.num_items, ptr(cand), self.cand_cap, ptr(count), ptr(thresh), ptr(inv_scale), ptr(self.item_params),
                 ptr(self.center), ptr(ex_ptr), ptr(excl[1]), k, ptr(ids[b0:b1]), ptr(sc[b0:b1]), ptr(cert[b0:b1]), s)
            note("rescore_end")
        if mark:
            self._debug = (count, thresh)
