"""Scratch: step time of the fused kernel at cfg-2 lattice size for several batch sizes (is the gather bound by
HBM or by what one SM keeps in flight?)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import warp_rnnt_b200 as w

dev = torch.device("cuda:0")
T, U, V = 150, 40, 28
out = {}
for N in (32, 64, 128, 148, 256, 296):
    sets = []
    nsets = max(2, int(700e6 // (8 * N * T * U * V)) + 1)
    for s in range(min(nsets, 8)):
        torch.manual_seed(s)
        xs = torch.log_softmax(torch.randn(N, T, U, V, device=dev), -1)
        ys = torch.randint(1, V, (N, U - 1), dtype=torch.int, device=dev)
        xn = torch.full((N,), T, dtype=torch.int, device=dev)
        yn = torch.full((N,), U - 1, dtype=torch.int, device=dev)
        sets.append((xs, ys, xn, yn))
    keep = [None] * len(sets)
    r = {}
    for mode in ("exact", "fast"):
        w.set_lse_mode(mode)
        for i in range(len(sets) + 3):
            keep[i % len(sets)] = w._C.rnnt_loss(*sets[i % len(sets)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(100):
            keep[i % len(sets)] = w._C.rnnt_loss(*sets[i % len(sets)])
        e1.record()
        torch.cuda.synchronize()
        r[mode] = e0.elapsed_time(e1) * 10.0   # us per step
    out[N] = r
    print(N, r, flush=True)
    del sets, keep
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/nsweep.json", "w"))
