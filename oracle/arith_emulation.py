"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): a CPU emulation of the engine's two operand arithmetics
(include/sce.h `sce_arith`), exact in everything but the tensor core's fp32 accumulation order —
products are accumulated in fp64 here. It pins the ERROR MODEL of the arithmetics without a GPU: what the planes can
represent, which partial products are formed, what is dropped.

  f16f8 : x ~= h + l,  h = fp16(x);  planes: h (fp16), h8 = e5m2(x), l8 = e5m2((x - h) * 2^11)
          a @ b ~= h_a @ h_b + (l8_a @ h8_b + h8_a @ l8_b) * 2^-11        (sce_ptx.cuh "fp16 + fp8 arithmetic")
  bf16x3: x ~= hi + lo, hi = bf16(x), lo = bf16(x - hi)
          a @ b ~= hi_a @ hi_b + lo_a @ hi_b + hi_a @ lo_b                 (sce_gemm.cuh)
"""
import torch

LO_SHIFT = 11            # kLoShift in sparse_coding_b200/csrc/sce_ptx.cuh
_S = float(1 << LO_SHIFT)


def planes_f16f8(a: torch.Tensor):
    """(h, h8, l8) as fp32 tensors holding exactly the values the device planes hold."""
    a = a.float()
    h = a.half().float()
    h8 = a.to(torch.float8_e5m2).float()
    l8 = ((a - h) * _S).to(torch.float8_e5m2).float()
    return h, h8, l8


def mm_f16f8(a: torch.Tensor, b: torch.Tensor, skip_a_residual: bool = False) -> torch.Tensor:
    """a [m,k] @ b [k,n] in the f16f8 arithmetic (fp64 accumulation). ``skip_a_residual`` drops the l8_a @ h8_b term,
    as the device does when a's residual plane is flagged all-zero."""
    ah, a8, al = planes_f16f8(a)
    bh, b8, bl = planes_f16f8(b)
    cross = a8.double() @ bl.double()
    if not skip_a_residual:
        cross = cross + al.double() @ b8.double()
    return ah.double() @ bh.double() + cross / _S


def mm_bf16x3(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = a.float(), b.float()
    ah = a.bfloat16().float()
    al = (a - ah).bfloat16().float()
    bh = b.bfloat16().float()
    bl = (b - bh).bfloat16().float()
    return ah.double() @ bh.double() + al.double() @ bh.double() + ah.double() @ bl.double()


def tied_step_emulated(mm, x, W, bias, alpha, pin_active=None):
    """Forward + backward GEMMs of the tied SAE in the given arithmetic, in the scaling the engine uses for f16f8 (the
    backward pass runs on the residual r, outputs are multiplied by 2/(B d)). W: unit-norm rows [n,d]. Returns
    (z, x_hat, dW); ``pin_active`` (bool [B,n]) pins the ReLU pattern."""
    B, d = x.shape
    z = mm(x, W.T.contiguous()).float() + bias
    active = (z > 0) if pin_active is None else pin_active
    c = (z * active).float()
    x_hat = mm(c, W).float()
    r = x_hat - x
    dz = ((mm(r, W.T.contiguous()).float() + alpha * d / 2) * active).float()
    dW = (mm(dz.T.contiguous(), x) + mm(c.T.contiguous(), r)) * (2.0 / (B * d))
    return z, x_hat, dW
