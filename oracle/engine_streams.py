"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (options_model_b200/).

Paired check of the network regression.  The reference draws its weight
initialisation, its DataLoader shuffle and its dropout masks from torch's global generator (om3:565-613,
om3gpu:740-798); a device kernel cannot consume that stream, so the engine uses counter-based streams instead (a Feistel
permutation per epoch, a hashed keep mask per (step, row, layer, unit) -- csrc/lsm_gnet.cu).  Those streams are pure
functions of (seed, epoch, step, row); this file restates them in numpy so that the torch fp32 restatement of the
reference's algorithm can be trained on the SAME initial weights, the SAME mini-batches and the SAME dropout masks as
the engine.  What is left between the two runs is arithmetic only (bf16 tensor-core operands in the two hidden layers
against fp32), which is what the tightened tolerances in tests/test_gpu_network.py measure.

`paired_gnet` follows om3gpu:700-830 (pass 1 rows, normalisation with the sample std, AdamW(1e-3, wd 1e-4), batch
min(8192, n), best-weights snapshot, decision pass with dropout left active) exactly like `lsm_global` +
`single_lsm_net_fit("gpu")` in lsm_oracle.py; the only difference is where the random bits come from.
"""
import copy
import math

import numpy as np

from .lsm_oracle import features_ref7, payoff

_U32 = np.uint32


def mix32(h):
    """The 32-bit finaliser both streams are built on."""
    with np.errstate(over="ignore"):
        h = np.asarray(h, dtype=_U32).copy()
        h ^= h >> _U32(16)
        h *= _U32(0x85EBCA6B)
        h ^= h >> _U32(13)
        h *= _U32(0xC2B2AE35)
        h ^= h >> _U32(16)
    return h


def perm_key(seed: int, epoch: int) -> int:
    return (((seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF) >> 32) + 0x632BE5AB * (epoch + 1) & 0xFFFFFFFF


def train_drop_key(seed: int, step: int) -> int:
    """`step` counts optimiser steps from 1, across epochs."""
    return ((seed & 0xFFFFFFFF) * 0x2545F491 + step * 0x9E3779B1) & 0xFFFFFFFF


def walk_drop_key(seed: int) -> int:
    return ((seed & 0xFFFFFFFF) * 0x2545F491 + 0x51ED270B) & 0xFFFFFFFF


def walk_row_id(j, t: int):
    """Row identifier of path j at date t in the decision pass."""
    with np.errstate(over="ignore"):
        return np.asarray(j, dtype=_U32) * _U32(0x01000193) + _U32(t)


def feistel_perm(n: int, key: int) -> np.ndarray:
    """src[i] = the row that position i of the epoch's order reads: a 4-round Feistel bijection on 2^(2 half) values,
    cycle-walked onto [0, n)."""
    if n <= 1:
        return np.arange(n, dtype=np.int64)
    bits = 1
    while (1 << bits) < n:
        bits += 1
    half = (bits + 1) // 2
    mask = _U32((1 << half) - 1)
    key = _U32(key)
    out = np.arange(n, dtype=np.uint64)
    todo = np.arange(n)
    with np.errstate(over="ignore"):
        while todo.size:
            i = out[todo]
            L = (i >> np.uint64(half)).astype(_U32) & mask
            R = i.astype(_U32) & mask
            for rd in range(4):
                f = mix32(R + key * _U32(2 * rd + 1) + _U32(0x9E3779B9) * _U32(rd + 1)) & mask
                L, R = R, L ^ f
            i = (L.astype(np.uint64) << np.uint64(half)) | R.astype(np.uint64)
            out[todo] = i
            todo = todo[i >= n]
    return out.astype(np.int64)


def keep_mask(key: int, row_id, layer: int, p: float, units: int = 128) -> np.ndarray:
    """[len(row_id), units] booleans: unit kept.  One hashed byte per unit against the threshold round(256 p); the realised
    keep probability is (256 - thr) / 256 and the survivors are scaled by its inverse (`keep_scale`)."""
    thr = min(max(int(round(p * 256.0)), 0), 255)
    r = np.asarray(row_id, dtype=_U32)
    if thr == 0:
        return np.ones((r.size, units), dtype=bool)
    keep = np.empty((r.size, units), dtype=bool)
    with np.errstate(over="ignore"):
        for c in range(units // 8):
            base = _U32(key) + r * _U32(0x9E3779B1) + _U32(((layer * 16 + c) * 0x7FEB352D) & 0xFFFFFFFF)
            h0, h1 = mix32(base), mix32(base ^ _U32(0x68E31DA4))
            for k in range(4):
                keep[:, 8 * c + k] = ((h0 >> _U32(8 * k)) & _U32(0xFF)) >= thr
                keep[:, 8 * c + 4 + k] = ((h1 >> _U32(8 * k)) & _U32(0xFF)) >= thr
    return keep


def keep_scale(p: float) -> float:
    thr = min(max(int(round(p * 256.0)), 0), 255)
    return 256.0 / (256 - thr) if thr else 1.0


def torch_default_init(seed: int, hidden: int = 128) -> np.ndarray:
    """SingleLSMNet(7, hidden, 3) (om3:85-103) under torch's default Linear initialisation, flattened in state_dict order
    (W1, b1, W2, b2, W3, b3, W4, b4) -- the layout of optmc_gnet_params.init_params."""
    import torch
    from torch import nn

    torch.manual_seed(seed)
    lins = [nn.Linear(7, hidden), nn.Linear(hidden, hidden), nn.Linear(hidden, hidden), nn.Linear(hidden, 1)]
    return np.concatenate([t.detach().numpy().reshape(-1) for l in lins for t in (l.weight, l.bias)]).astype(np.float32)


def paired_gnet(S, K, r, T, option_type, init_params, seed, epochs=25, lr=1e-3, weight_decay=1e-4, batch=8192, dropout=0.1,
                inference_dropout=True, min_delta=1e-6, hidden=128, log=None):
    """The torch-GPU file's algorithm (om3gpu:700-830; = lsm_global(..., single_lsm_net_fit("gpu"), target_ddof=1) with
    early stopping off) in fp32 torch on the CPU, driven by the engine's shuffle / dropout streams and `init_params`.
    Rows are laid out date-major, dates ascending, paths ascending -- the engine's row table order, which is what the
    permutation indexes.  -> (price, stats)."""
    import torch

    S = np.asarray(S, dtype=np.float64)
    N, M = S.shape[0] - 1, S.shape[1]
    dt = T / N
    term = payoff(S[-1], K, option_type).astype(np.float64)
    feats, targs = [], []
    for t in range(1, N):
        itm = payoff(S[t], K, option_type) > 0
        if np.any(itm):
            feats.append(features_ref7(S[t, itm], K, r, T, t * dt))
            targs.append(term[itm] * math.exp(-r * dt * (N - t)))
    if not feats:
        return float(term.mean() * math.exp(-r * dt * (N - 1))), None  # the reference never discounts date 1 -> 0 (om3:650)
    X_all, Y_all = np.vstack(feats), np.concatenate(targs)
    Y_mean, Y_std = Y_all.mean(), Y_all.std(ddof=1)
    if not Y_std > 0:
        Y_std = 1.0
    f_mean, f_std = X_all.mean(axis=0), X_all.std(axis=0)
    f_std[f_std == 0] = 1
    X = torch.from_numpy(((X_all - f_mean) / f_std).astype(np.float32))
    Y = torch.from_numpy(((Y_all - Y_mean) / Y_std).astype(np.float32))
    n = X.shape[0]

    p0 = torch.from_numpy(np.asarray(init_params, dtype=np.float32).copy())
    shapes = [(hidden, 7), (hidden,), (hidden, hidden), (hidden,), (hidden, hidden), (hidden,), (1, hidden), (1,)]
    params, o = [], 0
    for sh in shapes:
        k = int(np.prod(sh))
        params.append(p0[o:o + k].reshape(sh).clone().requires_grad_(True))
        o += k
    assert o == p0.numel()
    sc = keep_scale(dropout)

    def forward(x, masks):
        h = x
        for l in range(3):
            h = torch.relu(h @ params[2 * l].T + params[2 * l + 1])
            if masks is not None:
                h = h * (masks[l] * sc)
        return (h @ params[6].T + params[7]).reshape(-1)

    opt = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay)
    B = min(batch, n)
    best, best_sd, step = float("inf"), None, 0
    for ep in range(epochs):
        src = torch.from_numpy(feistel_perm(n, perm_key(seed, ep)))
        tot, nb = 0.0, 0
        for b0 in range(0, n, B):
            b1 = min(b0 + B, n)
            step += 1
            rows = np.arange(b0, b1, dtype=np.uint32)
            key = train_drop_key(seed, step)
            masks = [torch.from_numpy(keep_mask(key, rows, l, dropout, hidden).astype(np.float32)) for l in range(3)]
            idx = src[b0:b1]
            loss = torch.mean((forward(X[idx], masks) - Y[idx]) ** 2)
            opt.zero_grad()
            loss.backward()
            opt.step()
            tot += float(loss.detach())
            nb += 1
        avg = tot / nb
        if log is not None:
            log.append(avg)
        if avg < best - min_delta:
            best, best_sd = avg, [p.detach().clone() for p in params]
    if best_sd is not None:
        with torch.no_grad():
            for p, q in zip(params, best_sd):
                p.copy_(q)

    # decision pass (om3gpu:800-830): sticky mask, strict comparison, dropout left on
    disc = math.exp(-r * dt)
    cf = term.copy()
    exercised = np.zeros(M, dtype=bool)
    ex_count = np.zeros(N + 1, dtype=np.int64)
    wkey = walk_drop_key(seed)
    for t in range(N - 1, 0, -1):
        cf *= disc
        itm = (payoff(S[t], K, option_type) > 0) & (~exercised)
        if not np.any(itm):
            continue
        j = np.where(itm)[0]
        fn = torch.from_numpy(((features_ref7(S[t, itm], K, r, T, t * dt) - f_mean) / f_std).astype(np.float32))
        masks = None
        if inference_dropout:
            rid = walk_row_id(j, t)
            masks = [torch.from_numpy(keep_mask(wkey, rid, l, dropout, hidden).astype(np.float32)) for l in range(3)]
        with torch.no_grad():
            cont = forward(fn, masks).numpy().astype(np.float64) * Y_std + Y_mean
        pay = payoff(S[t, itm], K, option_type)
        ex = pay > cont
        cf[j[ex]] = pay[ex]
        exercised[j[ex]] = True
        ex_count[t] = int(ex.sum())
    stats = dict(n_rows=int(n), best_loss=best, ex_count=ex_count, Y_mean=float(Y_mean), Y_std=float(Y_std),
                 params=np.concatenate([p.detach().numpy().reshape(-1) for p in params]))
    return float(cf.mean()), stats
