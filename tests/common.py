"""Shared helpers for the tests: golden vectors, synthetic inputs (the reference's recipe)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))


def log_softmax(x, axis=-1):
    x = np.asarray(x, dtype=np.float64)
    m = x.max(axis=axis, keepdims=True)
    y = x - m
    return y - np.log(np.exp(y).sum(axis=axis, keepdims=True))


def golden_case(name):
    """-> dict(lp f32 (N,T,U,V) log-softmaxed, ys, xn, yn, costs, grads)."""
    r = GOLDEN[name]
    xs = np.asarray(r["xs"], dtype=np.float32)
    # the reference tests log_softmax in fp32 torch (test.py:42); fp64-then-round is within 1 ulp
    lp = log_softmax(xs).astype(np.float32)
    ys = np.asarray(r["ys"], dtype=np.int32).reshape(xs.shape[0], -1)
    costs = np.atleast_1d(np.asarray(r.get("expected_costs", r.get("expected_cost")), dtype=np.float64))
    return dict(lp=lp, ys=ys, xn=np.asarray(r["xn"], dtype=np.int32), yn=np.asarray(r["yn"], dtype=np.int32),
                costs=costs, grads=np.asarray(r["expected_grads"], dtype=np.float64))


def make_inputs(N, T, U, V, seed=0, random_lengths=False, blank=0, min_len=True):
    """Synthetic inputs following the reference's benchmark recipe (benchmark.py:11-27):
    randn -> log_softmax, labels in [1,V) (never blank when blank == 0), full or random lengths
    (random: benchmark.py:20-23, shifted so the max hits T / U-1)."""
    rng = np.random.RandomState(seed)
    xs = rng.randn(N, T, U, V).astype(np.float32)
    lp = log_softmax(xs).astype(np.float32)
    if V > 1:
        ys = rng.randint(0, V - 1, (N, max(U - 1, 0))).astype(np.int32)
        ys = np.where(ys >= blank, ys + 1, ys).astype(np.int32)     # skip the blank id
    else:
        ys = np.zeros((N, max(U - 1, 0)), dtype=np.int32)
    if random_lengths:
        xn = rng.randint(max(T // 2, 1), T + 1, (N,)).astype(np.int32)
        yn = rng.randint(U // 2, U, (N,)).astype(np.int32) if U > 1 else np.zeros(N, np.int32)
        xn = xn + T - xn.max()
        yn = yn + (U - 1) - yn.max()
    else:
        xn = np.full(N, T, dtype=np.int32)
        yn = np.full(N, U - 1, dtype=np.int32)
    return lp, ys, xn, yn


def to_compact(lp, ys, xn, yn):
    """Ragged concat as in the reference test (test.py:291-299)."""
    V = lp.shape[-1]
    xs_c = np.concatenate([lp[i, :xn[i], :yn[i] + 1].reshape(-1, V) for i in range(lp.shape[0])], axis=0)
    ys_c = np.concatenate([ys[i, :yn[i]] for i in range(ys.shape[0])], axis=0).astype(np.int32)
    return np.ascontiguousarray(xs_c), np.ascontiguousarray(ys_c)


def from_compact(flat, xn, yn, T, U):
    """(STU,V) -> padded (N,T,U,V) with zeros."""
    V = flat.shape[-1]
    N = len(xn)
    out = np.zeros((N, T, U, V), dtype=flat.dtype)
    o = 0
    for i in range(N):
        c = int(xn[i]) * (int(yn[i]) + 1)
        out[i, :xn[i], :yn[i] + 1] = flat[o:o + c].reshape(xn[i], yn[i] + 1, V)
        o += c
    return out
