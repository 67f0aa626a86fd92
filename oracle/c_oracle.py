"""ctypes wrapper of oracle/c/hnm_oracle.c, the plain-C restatement of the hot path (TEST INFRASTRUCTURE, see
oracle/__init__.py).  ``build()`` compiles it with gcc into oracle/_build/ (git-ignored; it travels to the GPU box
with the snapshot like every other built artefact); only tests/ use it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "hnm_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libhnm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (no fused multiply-adds: every fp32 product and sum is rounded as written)."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(OUT_DIR, exist_ok=True)
        cmd = ["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-fPIC",
               "-shared", "-Wall", "-Wextra", "-Werror", SRC, "-o", LIB, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(x, dtype) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


def norm_adj(edge_index, edge_weight, num_nodes: int):
    """(rowptr int64 [N+1], col int64 [nnz], val fp32 [nnz], dis fp32 [N]); src/models/lightgcn.py:92-112."""
    ei = _arr(edge_index, np.int64)
    m = ei.shape[1]
    row, col = np.ascontiguousarray(ei[0]), np.ascontiguousarray(ei[1])
    w = None if edge_weight is None else _arr(edge_weight, np.float32)
    rowptr = np.empty(num_nodes + 1, np.int64)
    out_col = np.empty(m + num_nodes, np.int64)
    out_val = np.empty(m + num_nodes, np.float32)
    dis = np.empty(num_nodes, np.float32)
    rc = load().hnm_oracle_norm_adj(_p(row), _p(col), _p(w), C.c_int64(m), C.c_int64(num_nodes), _p(rowptr), _p(out_col),
                                    _p(out_val), _p(dis))
    if rc:
        raise RuntimeError(f"hnm_oracle_norm_adj: {rc}")
    return rowptr, out_col, out_val, dis


def forward(weight, rowptr, col, val, num_users: int, num_layers: int, alphas):
    """(user_emb, item_emb) fp32; src/models/lightgcn.py:147-162."""
    e0 = _arr(weight, np.float32)
    n, d = e0.shape
    final = np.empty_like(e0)
    al = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64))
    rc = load().hnm_oracle_forward(_p(_arr(rowptr, np.int64)), _p(_arr(col, np.int64)), _p(_arr(val, np.float32)),
                                   C.c_int64(n), C.c_int32(d), C.c_int32(num_layers), _p(al), _p(e0), _p(final))
    if rc:
        raise RuntimeError(f"hnm_oracle_forward: {rc}")
    return final[:num_users], final[num_users:]


def topk_exact(user_emb, item_emb, user_ids, k: int, filter_items: Optional[Dict[int, set]] = None, item_bias=None):
    """(ids int64 [B, k], scores fp64 [B, k]) by (score desc, id asc); src/models/lightgcn.py:199-202,349-356.
    item_bias: MatrixFactorization's b_i (src/models/matrix_factorization.py:108-131), added to the exact score."""
    ue, ie = _arr(user_emb, np.float32), _arr(item_emb, np.float32)
    uids = _arr(user_ids, np.int64)
    b = uids.shape[0]
    ex_ptr = ex_items = None
    if filter_items is not None:
        ptr, items = [0], []
        for u in uids.tolist():
            items.extend(sorted(filter_items.get(int(u), ())))
            ptr.append(len(items))
        ex_ptr, ex_items = np.asarray(ptr, np.int64), np.asarray(items + [0], np.int64)
    ids = np.empty((b, k), np.int64)
    sc = np.empty((b, k), np.float64)
    rc = load().hnm_oracle_topk_exact(_p(ue), _p(ie), _p(uids), C.c_int64(b), C.c_int64(ie.shape[0]), C.c_int32(ie.shape[1]),
                                      C.c_int32(k), _p(ex_ptr), _p(ex_items),
                                      _p(None if item_bias is None else _arr(item_bias, np.float32).ravel()), _p(ids), _p(sc))
    if rc == -3:
        raise RuntimeError("selected index k out of range")
    if rc:
        raise RuntimeError(f"hnm_oracle_topk_exact: {rc}")
    return ids, sc


def ncf_forward(state: Dict[str, object], user_ids, item_ids) -> np.ndarray:
    """fp32 logits [B]; ``state`` uses the reference's state_dict keys; src/models/neural_cf.py:125-139."""
    gu, gi = _arr(state["gmf_user_embedding.weight"], np.float32), _arr(state["gmf_item_embedding.weight"], np.float32)
    mu, mi = _arr(state["mlp_user_embedding.weight"], np.float32), _arr(state["mlp_item_embedding.weight"], np.float32)
    ws, bs, dims, i = [], [], [2 * mu.shape[1]], 0
    while f"mlp_layers.{i}.weight" in state:                      # Linear modules sit at 0, 3, 6, ... (:85-90)
        w = _arr(state[f"mlp_layers.{i}.weight"], np.float32)
        ws.append(w.ravel())
        bs.append(_arr(state[f"mlp_layers.{i}.bias"], np.float32))
        dims.append(w.shape[0])
        i += 3
    weights = np.ascontiguousarray(np.concatenate(ws))
    biases = np.ascontiguousarray(np.concatenate(bs))
    dims_a = np.asarray(dims, np.int32)
    pw = _arr(state["prediction_layer.weight"], np.float32).ravel()
    pb = float(_arr(state["prediction_layer.bias"], np.float32).ravel()[0])
    u, it = _arr(user_ids, np.int64), _arr(item_ids, np.int64)
    out = np.empty(u.shape[0], np.float32)
    rc = load().hnm_oracle_ncf_forward(_p(gu), _p(gi), _p(mu), _p(mi), C.c_int32(gu.shape[1]), C.c_int32(mu.shape[1]),
                                       C.c_int32(len(ws)), _p(dims_a), _p(weights), _p(biases), _p(pw), C.c_float(pb),
                                       _p(u), _p(it), C.c_int64(u.shape[0]), _p(out))
    if rc:
        raise RuntimeError(f"hnm_oracle_ncf_forward: {rc}")
    return out
