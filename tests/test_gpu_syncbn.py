"""GPU: SyncBatchNorm kernels against the fixtures / oracle (single process; the N-replica
collective logic is covered by tests/test_sharded_gloo.py on CPU and by bench --gpus N)."""
import json

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden_names, load_golden
from helpers import relmax
from oracle import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("bn_"))
def test_kernels_reproduce_n_replica_formula(cuda_lib, name):
    """stats on each shard + sum == the reference's reduce; finalize(clamp) + apply == fixture."""
    from eegan_b200.sync_batchnorm.batchnorm import CudaBNOps as K
    g = load_golden(name)
    kw = json.loads(str(g["recipe"]))
    C = kw["C"]
    x = torch.from_numpy(g["x"]).cuda()
    shards = [s.contiguous().reshape(s.shape[0], C, -1) for s in x.chunk(kw["shards"], 0)]
    tot = torch.zeros(2 * C, device="cuda")
    for s in shards:
        buf = torch.empty(2 * C, device="cuda")
        K.stats(s, buf)
        tot += buf
    count = sum(s.shape[0] * s.shape[2] for s in shards)
    mean, inv_std = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    K.finalize(tot, C, count, None, 1e-5, 0.1, 1, mean, inv_std, rm, rv)
    np.testing.assert_allclose(mean.cpu().numpy(), g["mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(inv_std.cpu().numpy(), g["inv_std"], rtol=1e-5)
    np.testing.assert_allclose(rm.cpu().numpy(), g["running_mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rv.cpu().numpy(), g["running_var"], rtol=1e-5)
    w, b = torch.from_numpy(g["weight"]).cuda(), torch.from_numpy(g["bias"]).cuda()
    outs = []
    for s in shards:
        y = torch.empty_like(s)
        K.apply(s, mean, inv_std, w, b, y)
        outs.append(y)
    got = torch.cat(outs, 0).reshape(x.shape).cpu().numpy()
    np.testing.assert_allclose(got, g["out"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape,affine", [((8, 32, 16, 16), True), ((4, 100, 7, 7), True), ((6, 256, 4, 4), False),
                                          ((2, 32, 64, 64), True)])
def test_module_single_replica_matches_batch_norm(cuda_lib, shape, affine):
    """One replica: the reference is F.batch_norm (batchnorm.py:50-53) — forward, running
    statistics, backward and eval mode."""
    from eegan_b200.sync_batchnorm import SynchronizedBatchNorm2d
    c = cases.bn_case(shape[0], shape[1], shape[2])
    x = c["x"][..., : shape[3]].contiguous() if shape[3] != shape[2] else c["x"]
    bn = SynchronizedBatchNorm2d(shape[1], affine=affine).cuda()
    ref = torch.nn.BatchNorm2d(shape[1], affine=affine).double()
    if affine:
        with torch.no_grad():
            bn.weight.copy_(c["weight"]); bn.bias.copy_(c["bias"])
            ref.weight.copy_(c["weight"]); ref.bias.copy_(c["bias"])
    xg = x.cuda().requires_grad_()
    xr = x.double().requires_grad_()
    y, yr = bn(xg), ref(xr)
    assert relmax(y.detach().cpu(), yr.detach()) <= 1e-5
    gy = torch.randn(y.shape, generator=cases._gen(8))
    (y * gy.cuda()).sum().backward()
    (yr * gy.double()).sum().backward()
    assert relmax(xg.grad.cpu(), xr.grad) <= 1e-4
    if affine:
        assert relmax(bn.weight.grad.cpu(), ref.weight.grad) <= 1e-4
        assert relmax(bn.bias.grad.cpu(), ref.bias.grad) <= 1e-4
    assert relmax(bn.running_mean.cpu(), ref.running_mean) <= 1e-5
    assert relmax(bn.running_var.cpu(), ref.running_var) <= 1e-5
    bn.eval(); ref.eval()
    assert relmax(bn(x.cuda()).cpu(), ref(x.double())) <= 1e-5
    # eval mode is differentiable, as the reference's F.batch_norm branch is (batchnorm.py:50-53)
    xg2, xr2 = x.cuda().requires_grad_(), x.double().requires_grad_()
    bn.zero_grad(set_to_none=True); ref.zero_grad(set_to_none=True)
    y2, yr2 = bn(xg2), ref(xr2)
    assert y2.grad_fn is not None
    (y2 * gy.cuda()).sum().backward()
    (yr2 * gy.double()).sum().backward()
    assert relmax(xg2.grad.cpu(), xr2.grad) <= 1e-5
    if affine:
        assert relmax(bn.weight.grad.cpu(), ref.weight.grad) <= 1e-4
        assert relmax(bn.bias.grad.cpu(), ref.bias.grad) <= 1e-4


@pytest.mark.parametrize("shape", [(32, 256, 16, 16), (8, 64, 32, 32), (3, 32, 5, 7), (2, 40, 128, 128)])
def test_single_replica_fused_calls_match_the_separate_steps(cuda_lib, shape):
    """eegan_syncbn_fwd_fused / _bwd_fused (one launch per direction on small maps, the stats/finalize/apply sequence on
    large ones) against the separate entry points on the same data: same arithmetic, different launch structure."""
    from eegan_b200.sync_batchnorm.batchnorm import CudaBNOps as K
    N, C, H, W = shape
    g = cases._gen(N * C + H)
    x = (torch.randn(N, C, H * W, generator=g) * 1.3 + 0.2).cuda()
    dy = torch.randn(N, C, H * W, generator=g).cuda()
    w, b = (torch.rand(C, generator=g) + 0.5).cuda(), torch.randn(C, generator=g).cuda()
    # separate steps
    buf = torch.empty(2 * C, device="cuda"); K.stats(x, buf)
    mean, inv = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    K.finalize(buf, C, N * H * W, None, 1e-5, 0.1, 0, mean, inv, rm, rv)
    y = torch.empty_like(x); K.apply(x, mean, inv, w, b, y)
    red = torch.empty(2 * C, device="cuda"); K.bwd_reduce(x, dy, mean, inv, red)
    dx = torch.empty_like(x); K.bwd_apply(x, dy, mean, inv, w, red, N * H * W, None, 1e-5, 0, dx)
    # fused
    work = torch.empty(4 * C, device="cuda")
    rm2, rv2 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    y2 = torch.empty_like(x); K.fwd_fused(x, w, b, 1e-5, 0.1, rm2, rv2, y2, work)
    red2 = torch.empty(2 * C, device="cuda"); dx2 = torch.empty_like(x)
    K.bwd_fused(x, dy, work, w, 1e-5, dx2, red2)
    assert relmax(work[2 * C:3 * C], mean) <= 1e-5 and relmax(work[3 * C:], inv) <= 1e-5
    assert float((y2 - y).abs().max()) <= 1e-5 and relmax(dx2, dx) <= 1e-4
    assert relmax(red2, red) <= 1e-4 and relmax(rm2, rm) <= 1e-5 and relmax(rv2, rv) <= 1e-5
