"""Scratch timing of the compact-layout forward (+backward) at a cfg-2 sized ragged batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import warp_rnnt_b200 as w

dev = torch.device("cuda:0")
N, T, U, V = 128, 150, 40, 28
rng = np.random.default_rng(0)
xn = rng.integers(T // 2, T + 1, N).astype(np.int32); xn[0] = T
yn = rng.integers(U // 2, U, N).astype(np.int32); yn[1] = U - 1
STU = int((xn.astype(np.int64) * (yn + 1)).sum())
sets = []
for s in range(4):
    torch.manual_seed(s)
    xs = torch.log_softmax(torch.randn(STU, V, device=dev), -1)
    ys = torch.randint(1, V, (int(yn.sum()),), dtype=torch.int, device=dev)
    sets.append((xs, ys, torch.from_numpy(xn).to(dev), torch.from_numpy(yn).to(dev)))
cum = torch.cumsum(torch.from_numpy(xn.astype(np.int64) * (yn + 1)), 0).int().to(dev)


def run(i, bwd):
    a = sets[i % 4]
    costs, pg, loc = w._C.rnnt_loss_compact(*a, 0, 0.0, True)
    if bwd:
        w._C.rnnt_loss_compact_backward(torch.ones_like(costs), pg, cum, loc, V, 0)


for bwd in (False, True):
    for i in range(5):
        run(i, bwd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        run(i, bwd)
    e1.record()
    torch.cuda.synchronize()
    print("path=%s %s: %.1f us/step (STU=%d)" % (os.environ.get("RNNT_B200_PATH", "auto"), "fwd+bwd" if bwd else "fwd", e0.elapsed_time(e1) / 40 * 1e3, STU))
