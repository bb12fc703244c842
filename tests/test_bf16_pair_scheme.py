"""The operand scheme of gag_tc_bwd.cu, checked as arithmetic on the CPU (no kernel involved).

The one-pass backward of GlobalAttentionGeneral stores every operand as a PAIR of bf16 values, hi = bf16(v) and
lo = bf16(v - hi), and contracts with three tensor-core products per K-step: lo*hi + hi*lo + hi*hi, accumulated in fp32
(the lo*lo term, ~2^-18 of the product, is dropped).  This file emulates that arithmetic with torch on the CPU for the four
contraction shapes of the kernel and holds it to the bound DESIGN.md 4.3 states: ~1e-5 of the result's maximum, an order of
magnitude inside the 1e-4 relative-to-max the gradients are tested to.  It also shows why one bf16 per operand (a single
pass) is not an option, and that the pair carries 16 significant bits whatever the magnitude (no scaling needed, unlike
the fp16 pairs of the pair grid)."""
import pytest
import torch


def split(v):
    hi = v.to(torch.bfloat16).to(torch.float32)
    lo = (v - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def pair_matmul(a, b):
    """a [M, K] @ b [K, N] the way the kernel contracts: three products, fp32 accumulation"""
    ah, al = split(a)
    bh, bl = split(b)
    return al @ bh + ah @ bl + ah @ bh


def relmax(x, ref):
    return float((x.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M,K,N,what", [(128, 128, 18, "(1) dP = d_out^T value: K = channels"),
                                         (64, 4096, 18, "(2)/(4) dV, dK: K = pixels of a group"),
                                         (128, 18, 128, "(3) dX = ds key^T: K = words")])
def test_three_products_of_bf16_pairs_are_fp32_class(M, K, N, what):
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(K, N, generator=g) * (1.0 if K > 64 else 0.3)
    ref = a.double() @ b.double()
    err3 = relmax(pair_matmul(a, b), ref)
    err1 = relmax(a.to(torch.bfloat16).to(torch.float32) @ b.to(torch.bfloat16).to(torch.float32), ref)
    assert err3 <= 2e-5, (what, err3)
    assert err1 >= 20 * err3 and err1 > 1e-4, (what, err1, err3)  # one bf16 per operand: far outside the gradients' tolerance


@pytest.mark.parametrize("scale", [1e-30, 1e-6, 1.0, 3e4, 1e30])
def test_the_pair_keeps_16_bits_at_any_magnitude(scale):
    """bf16 has fp32's exponent range: hi + lo reproduces v to ~2^-17 relative whether v is 1e-30 or 1e30 — which is why the
    GlobalAttentionGeneral kernels need no per-tensor scale (upstream gradients can have any magnitude)."""
    g = torch.Generator().manual_seed(7)
    v = torch.randn(4096, generator=g) * scale
    hi, lo = split(v)
    rel = ((hi.double() + lo.double() - v.double()).abs() / v.double().abs().clamp(min=1e-300)).max()
    assert float(rel) <= 2.0 ** -16


def test_softmax_backward_chain_within_tolerance():
    """dP -> ds = p (dP - sum p dP) -> dX with every contraction on bf16 pairs and p re-read as hi + lo (as the pixel warps do):
    the chain's error stays ~1e-5 of the maximum."""
    g = torch.Generator().manual_seed(11)
    Q, C, T = 512, 128, 18
    d_out = torch.randn(C, Q, generator=g)
    value = torch.randn(C, T, generator=g)
    key = torch.randn(C, T, generator=g) * C ** -0.5
    p = torch.softmax(torch.randn(Q, T, generator=g) * 2.0, dim=1)
    dP = pair_matmul(d_out.t().contiguous(), value)
    ph, pl = split(p)
    pr = ph + pl
    ds = pr * (dP - (pr * dP).sum(dim=1, keepdim=True))
    dX = pair_matmul(ds, key.t().contiguous())
    dPr = d_out.double().t() @ value.double()
    dsr = p.double() * (dPr - (p.double() * dPr).sum(dim=1, keepdim=True))
    dXr = dsr @ key.double().t()
    assert relmax(dX, dXr) <= 3e-5
