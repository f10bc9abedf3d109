"""Scratch: where does exact mode differ from the compiled reference?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import warp_rnnt_b200 as w
from oracle import build_ref, oracle
from tests.common import make_inputs, to_compact
ref = build_ref.load()
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for shape in [(3, 7, 5, 6, True, 0, 0.0), (3, 150, 40, 28, True, 0, 0.01), (2, 70, 129, 7, False, 6, 0.0)]:
    N, T, U, V, rl, blank, lam = shape
    lp, ys, xn, yn = make_inputs(N, T, U, V, seed=21 + T, random_lengths=rl, blank=blank)
    args = (cu(lp), cu(ys), cu(xn), cu(yn))
    cr, gr = ref.rnnt_loss(*args, blank=blank, fastemit_lambda=lam)
    for mode in ("exact", "fast"):
        w.set_lse_mode(mode)
        cm, gm = w._C.rnnt_loss(*args, blank=blank, fastemit_lambda=lam)
        d = (gm - gr).abs()
        nz = (gm != gr).nonzero()
        print(shape, mode, "cost equal", torch.equal(cm, cr), (cm - cr).abs().max().item(), "grad maxdiff", d.max().item(),
              "n mismatched", nz.shape[0], "first", nz[:5].tolist())
    c64, g64 = oracle.dense(lp, ys, xn, yn, blank=blank, fastemit_lambda=lam)
    print("   ref vs f64: cost rel", np.abs(cr.cpu().numpy() - c64).max() / np.abs(c64).max(), "grad", np.abs(gr.cpu().numpy() - g64).max(),
          " fast vs f64 grad", np.abs(gm.cpu().numpy() - g64).max())
    # gathered
    idx = np.full((N, T, U, 2), blank, dtype=np.int64); idx[:, :, :U - 1, 1] = ys[:, None, :]
    g = np.ascontiguousarray(np.take_along_axis(lp, idx, axis=3))
    gargs = (cu(g), cu(ys), cu(xn), cu(yn))
    cr2, gr2 = ref.rnnt_loss(*gargs, blank=-1, fastemit_lambda=lam)
    w.set_lse_mode("exact")
    cm2, gm2 = w._C.rnnt_loss(*gargs, blank=-1, fastemit_lambda=lam)
    nz = (gm2 != gr2).nonzero()
    print("   gathered exact: cost eq", torch.equal(cm2, cr2), "mismatches", nz.shape[0], nz[:6].tolist(),
          [(gm2[tuple(i)].item(), gr2[tuple(i)].item()) for i in nz[:6]], "xn", xn, "yn", yn)
    # compact
    xs_c, ys_c = to_compact(lp, ys, xn, yn)
    cargs = (cu(xs_c), cu(ys_c), cu(xn), cu(yn))
    cr, gr, lr = ref.rnnt_loss_compact(*cargs, blank=blank, fastemit_lambda=lam)
    w.set_lse_mode("exact")
    cm, gm, lm = w._C.rnnt_loss_compact(*cargs, blank=blank, fastemit_lambda=lam)
    nz = (gm != gr).nonzero()
    print("   compact exact: cost eq", torch.equal(cm, cr), "loc eq", torch.equal(lm, lr), "grad mismatches", nz.shape[0], nz[:5].tolist(),
          (gm - gr).abs().max().item())
