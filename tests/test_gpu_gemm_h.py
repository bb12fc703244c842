"""GPU: the tcgen05 half-pair ("3xFP16") contraction engine (eegan_gemm_f16x3, gemm_h.cu) against float64
matmul: ragged extents (TMA zero-fill), padded pitches, batches, operand scales, and the two-accumulator form.
The bar is the same fp32-class accuracy as the 3xTF32 engine: max |err| <= 4e-6 * (|A| |B|) row-sum scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_case(L, M, N, K, batch, seed=0, sa=64.0, sb=64.0, dual=0, amp_a=1.0, amp_b=1.0):
    from eegan_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(seed)
    pad8 = lambda v: (v + 7) // 8 * 8
    A = torch.randn(batch, M, K, generator=g) * amp_a
    B = torch.randn(batch, N, K, generator=g) * amp_b
    lda = pad8(M) + 8
    Ab = torch.zeros(batch, K, lda)
    Ab[:, :, :M] = A.transpose(1, 2)
    ldb = pad8(K)
    Bb = torch.zeros(batch, N, ldb)
    Bb[:, :, :K] = B
    Ad, Bd = Ab.cuda(), Bb.cuda()
    ldc = N + 3
    C = torch.full((batch, M, ldc), -7.0, device="cuda")
    ws = torch.empty((2 if dual else 1) * 2 * (Ad.numel() * 2 + Bd.numel() * 2 + 512) + 512, dtype=torch.uint8, device="cuda")
    rc = L.eegan_gemm_f16x3(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(C), M, N, K, lda, ldb, ldc, Ad.stride(0), Bd.stride(0),
                            C.stride(0), batch, sa, sb, dual, _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "gemm_f16x3")
    torch.cuda.synchronize()
    ref = torch.bmm(A.double(), B.double().transpose(1, 2)) * (2.0 if dual else 1.0)
    got = C[:, :, :N].cpu().double()
    assert torch.all(C[:, :, N:] == -7.0), "wrote outside the N extent"
    scale = torch.bmm(A.double().abs(), B.double().abs().transpose(1, 2)).max().item() * (2.0 if dual else 1.0)
    err = (got - ref).abs().max().item()
    return err, scale


@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 32, 1), (128, 128, 64, 1), (289, 552, 256, 3), (256, 289, 540, 2),
                                         (100, 40, 289, 2), (640, 256, 289, 5)])
def test_f16x3_matches_fp64(cuda_lib, M, N, K, batch):
    err, scale = run_case(cuda_lib, M, N, K, batch, seed=M + N + K)
    assert err <= 4e-6 * scale, "max err %.3e vs scale %.3e" % (err, scale)


@pytest.mark.parametrize("M,N,K,batch", [(256, 289, 640, 3), (130, 70, 96, 2)])
def test_f16x3_dual_accumulators(cuda_lib, M, N, K, batch):
    err, scale = run_case(cuda_lib, M, N, K, batch, seed=7, dual=1)
    assert err <= 4e-6 * scale, "max err %.3e vs scale %.3e" % (err, scale)


def test_f16x3_scales_cover_small_and_large_operands(cuda_lib):
    # operands of magnitude 1e-3 and 30 stored with matching power-of-two scales keep fp32-class accuracy
    err, scale = run_case(cuda_lib, 289, 192, 256, 2, seed=3, sa=2.0 ** 16, sb=2.0 ** 7, amp_a=1e-3, amp_b=30.0)
    assert err <= 4e-6 * scale, "max err %.3e vs scale %.3e" % (err, scale)
