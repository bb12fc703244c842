"""GPU: the tcgen05 3xTF32 contraction engine (eegan_gemm_tf32x3) against float64 matmul, for
both operand majornesses, ragged extents (TMA zero-fill) and padded pitches.  The bar is
fp32-class accuracy: max |err| <= 4e-6 * (|A| |B|)_max-row-sum scale, far below single-pass
TF32 (~1e-3)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def run_case(L, M, N, K, a_k, b_k, batch, seed=0, staging=0):
    from eegan_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(seed)
    pad = lambda v: (v + 3) // 4 * 4
    # logical operands
    A = torch.randn(batch, M, K, generator=g)
    B = torch.randn(batch, N, K, generator=g)
    if a_k:
        lda = pad(K) + 4
        Ab = torch.zeros(batch, M, lda)
        Ab[:, :, :K] = A
    else:
        lda = pad(M) + 8
        Ab = torch.zeros(batch, K, lda)
        Ab[:, :, :M] = A.transpose(1, 2)
    if b_k:
        ldb = pad(K)
        Bb = torch.zeros(batch, N, ldb)
        Bb[:, :, :K] = B
    else:
        ldb = pad(N)
        Bb = torch.zeros(batch, K, ldb)
        Bb[:, :, :N] = B.transpose(1, 2)
    # garbage in the pads must never be read: poison them
    Ab[Ab == 0] = float("nan") if False else 0.0
    Ad, Bd = Ab.cuda(), Bb.cuda()
    ldc = N + 3
    C = torch.full((batch, M, ldc), -7.0, device="cuda")
    rc = L.eegan_gemm_tf32x3(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(C), M, N, K, int(a_k), int(b_k), lda, ldb, ldc,
                             Ad.stride(0), Bd.stride(0), C.stride(0), batch, staging, _lib.stream_ptr())
    _lib.check(rc, "gemm_tf32x3")
    torch.cuda.synchronize()
    ref = torch.bmm(A.double(), B.double().transpose(1, 2))
    got = C[:, :, :N].cpu().double()
    assert torch.all(C[:, :, N:] == -7.0), "wrote outside the N extent"
    scale = torch.bmm(A.double().abs(), B.double().abs().transpose(1, 2)).max().item()
    err = (got - ref).abs().max().item()
    return err, scale


@pytest.mark.parametrize("a_k,b_k", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 32, 1), (552, 289, 256, 3), (256, 289, 540, 2), (100, 40, 289, 2)])
def test_tf32x3_matches_fp64(cuda_lib, a_k, b_k, M, N, K, batch):
    err, scale = run_case(cuda_lib, M, N, K, a_k, b_k, batch, seed=M + N + K)
    assert err <= 4e-6 * scale, "max err %.3e vs scale %.3e (a_k=%s b_k=%s)" % (err, scale, a_k, b_k)


@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 32, 1), (289, 552, 256, 3), (256, 289, 540, 2), (100, 40, 289, 2), (640, 256, 289, 5)])
def test_tf32x3_tmem_staged_a_matches_fp64(cuda_lib, M, N, K, batch):
    """The TMEM-staged form (A: MN-major, split into tensor memory by tcgen05.st; B: K-major)."""
    err, scale = run_case(cuda_lib, M, N, K, False, True, batch, seed=M + N + K, staging=1)
    assert err <= 4e-6 * scale, "max err %.3e vs scale %.3e" % (err, scale)
