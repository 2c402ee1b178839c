"""CPU-side pin of the operand arithmetics' error model (oracle/arith_emulation.py): the bars the GPU parity tests
hold the engine to must already hold for an exact emulation of the planes and partial products — if these fail, no
kernel could pass; if they pass and a GPU test fails, the kernel (not the arithmetic) is wrong."""
import pytest
import torch

from oracle import arith_emulation as A

REL = 1e-4


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def _problem(d, n, B, seed, outliers=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, d, generator=g)
    if outliers:
        x[:, :4] *= 200.0          # massive-activation dimensions, as in LM residual streams
    E = torch.randn(n, d, generator=g)
    W = E / E.norm(dim=1, keepdim=True)
    bias = 0.02 * torch.randn(n, generator=g)
    return x, W, bias


def _fp64(x, W, bias, alpha):
    B, d = x.shape
    xd, Wd, bd = x.double(), W.double(), bias.double()
    z = xd @ Wd.T + bd
    act = z > 0
    c = z.clamp(min=0)
    xh = c @ Wd
    r = xh - xd
    dz = (r @ Wd.T * (2 / (B * d)) + alpha / B) * act
    dW = dz.T @ xd + c.T @ (r * (2 / (B * d)))
    return z, act, xh, dW


@pytest.mark.parametrize("outliers", [False, True])
@pytest.mark.parametrize("name,mm,bound", [("f16f8", A.mm_f16f8, 4e-5), ("bf16x3", A.mm_bf16x3, 1e-5)])
def test_forward_and_pinned_gradient_error(name, mm, bound, outliers):
    """x_hat within the 1e-4 bar with margin; with the ReLU pattern pinned the weight gradient is as accurate as the
    forward pass (un-pinned comparisons are dominated by coefficients at the kink, see tests/test_engine_gpu.py)."""
    x, W, bias = _problem(128, 512, 256, 0, outliers)
    alpha = 1e-2
    z64, act, xh64, dW64 = _fp64(x, W, bias, alpha)
    z, xh, dW = A.tied_step_emulated(mm, x, W, bias, alpha, pin_active=act)
    assert rel(z, z64) <= bound, (name, rel(z, z64))
    assert rel(xh, xh64) <= bound < REL, (name, rel(xh, xh64))
    assert rel(dW, dW64) <= bound, (name, rel(dW, dW64))


def test_f16f8_planes_and_exactness_flag():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(64, 96, generator=g) * 3.0
    h, h8, l8 = A.planes_f16f8(a)
    # the planes reconstruct a to 2^-11 * 2^-3 relative (fp16 plane + 3-bit residual)
    assert float(((h + l8 / 2048.0) - a).abs().max() / a.abs().max()) <= 2.0 ** -13
    # fp16-representable values have an all-zero residual plane: dropping its term changes nothing
    ae = a.half().float()
    assert float(A.planes_f16f8(ae)[2].abs().max()) == 0.0
    b = torch.randn(96, 40, generator=g)
    assert torch.equal(A.mm_f16f8(ae, b), A.mm_f16f8(ae, b, skip_a_residual=True))
    # ... and it matters otherwise (the term is ~2^-12 of the product)
    d = rel(A.mm_f16f8(a, b, skip_a_residual=True), a.double() @ b.double())
    assert 2e-5 < d < 1e-3, d


def test_f16f8_range_limits():
    """What the range monitor (sce_input_absmax) guards: beyond 65504 the fp16 plane overflows to inf (visible in the
    losses); far below 1e-4 the fp16 plane goes subnormal and relative precision degrades (bf16x3 keeps fp32 range)."""
    b = torch.randn(32, 8, generator=torch.Generator().manual_seed(2))
    big = torch.full((4, 32), 7.0e4)
    assert not torch.isfinite(A.mm_f16f8(big, b)).all()
    assert torch.isfinite(A.mm_bf16x3(big, b)).all()
    tiny = torch.randn(64, 32, generator=torch.Generator().manual_seed(3)) * 1e-7
    ref = tiny.double() @ b.double()
    assert rel(A.mm_bf16x3(tiny, b), ref) < 1e-5
    assert rel(A.mm_f16f8(tiny, b), ref) > 1e-4
