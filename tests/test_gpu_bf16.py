"""GPU: bfloat16 I/O of the dense path: bf16 log_probs in, bf16 gradient out, fp32 costs / alpha /
beta.  Checked against the fp64 oracle run on the SAME (bf16-representable) inputs:
  costs    |d|/|ref| <= 1e-5   (identical inputs, fp32 arithmetic, like the fp32 path)
  grads    |d| <= 2^-8 |g| + gtol(T,U)   (one bf16 rounding of the final value: 2^-9 relative, + the fp32 path's noise)
and against this library's fp32 path on the same inputs: the bf16 gradient must be the fp32 gradient rounded to bf16
(bit for bit, exact mode)."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.common import make_inputs

pytestmark = pytest.mark.gpu


def gtol(T, U):
    return 2e-5 + 6e-6 * (T + U)


@pytest.fixture(scope="module")
def w():
    import warp_rnnt_b200
    return warp_rnnt_b200


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# fused kernel (TMA rows / LDG gather / several CTAs per lattice) and general path (serial; the last one pipelined:
# its emit is long against its wavefront, see api.cu)
SHAPES = [(3, 20, 9, 8, True, 0, 0.0), (4, 150, 40, 28, True, 0, 0.0), (2, 33, 34, 5, True, 2, 0.25),
          (3, 40, 12, 1000, False, 0, 0.0), (2, 700, 40, 6, True, 0, 0.0), (2, 300, 50, 50, True, 0, 0.1),
          (32, 300, 70, 1200, True, 0, 0.0)]


@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("shape", SHAPES)
def test_bf16_dense_vs_oracle_and_fp32_path(w, shape, mode):
    N, T, U, V, rl, blank, lam = shape
    lp, ys, xn, yn = make_inputs(N, T, U, V, seed=3 + T, random_lengths=rl, blank=blank)
    xb = cu(lp).to(torch.bfloat16)                                  # the inputs ARE bf16 values
    lp_r = xb.float().cpu().numpy()
    args = (cu(ys), cu(xn), cu(yn))
    w.set_lse_mode(mode)
    try:
        costs, grads, loss = w._C.rnnt_loss_fused(xb, *args, blank, lam, None, True, 0)
        assert grads.dtype == torch.bfloat16 and costs.dtype == torch.float32
        c32, g32 = w._C.rnnt_loss(xb.float(), *args, blank=blank, fastemit_lambda=lam)
        assert torch.equal(costs, c32)                              # same inputs, same fp32 arithmetic
        assert torch.equal(grads, g32.to(torch.bfloat16))           # = the fp32 gradient, rounded once
        np.testing.assert_allclose(loss.item(), c32.sum().item(), rtol=2e-6)
        if N * T * U * V <= 2e7:
            c0, g0 = oracle.dense(lp_r, ys, xn, yn, blank=blank, fastemit_lambda=lam)
            np.testing.assert_allclose(costs.cpu().numpy(), c0, rtol=1e-5)
            err = np.abs(grads.float().cpu().numpy() - g0) - np.abs(g0) * 2.0 ** -8
            assert err.max() <= 2 * gtol(T, U), err.max()
    finally:
        w.set_lse_mode("auto")


def test_bf16_python_api_autograd(w):
    N, T, U, V = 4, 60, 21, 32
    lp, ys, xn, yn = make_inputs(N, T, U, V, seed=9, random_lengths=True)
    xb = cu(lp).to(torch.bfloat16)
    args = (cu(ys), cu(xn), cu(yn))
    for reduction, avg, scale in (("mean", True, 1.0), ("sum", False, 2.0), ("none", False, 1.0)):
        x = xb.clone().requires_grad_(True)
        loss = w.rnnt_loss(x, *args, average_frames=avg, reduction=reduction)
        assert loss.dtype == torch.float32
        (loss.sum() * scale).backward()
        assert x.grad.dtype == torch.bfloat16
        loss0, g0 = oracle.rnnt_loss(xb.float().cpu().numpy(), ys, xn, yn, avg, reduction)
        np.testing.assert_allclose(loss.detach().cpu().numpy(), loss0, rtol=1e-5)
        g0 = g0 * scale
        err = np.abs(x.grad.float().cpu().numpy() - g0) - np.abs(g0) * 2.0 ** -7   # two roundings when rescaled
        assert err.max() <= 2 * gtol(T, U), err.max()
    with pytest.raises(AssertionError):
        w.rnnt_loss(xb, *args, gather=True)
