"""GPU: edge cases of the pair grid the reference's semantics define — single pair, one-word
captions, T_max = 32, non-CUB shapes, all-same-class batches (every off-diagonal cell -inf),
and a batch larger than one GEMM tile in every dimension."""
import pytest
import torch

from helpers import finite_close, relmax
from oracle import cases
from oracle import damsm_oracle as O

pytestmark = pytest.mark.gpu


def _run_both(c, B, w0=1.0, w1=1.0, use_port=False):
    import eegan_b200 as E
    img = c["img"].cuda().requires_grad_()
    words = c["words"].cuda().requires_grad_()
    l0, l1, att = E.words_loss(img, words, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
    (w0 * l0 + w1 * l1).backward()
    if use_port:
        io, wo = c["img"].clone().requires_grad_(), c["words"].clone().requires_grad_()
        o0, o1, oatt = O.port_words_loss(io, wo, c["labels"], c["cap_lens"], c["class_ids"], B)
    else:
        io, wo = c["img"].double().requires_grad_(), c["words"].double().requires_grad_()
        o0, o1, oatt, _ = O.dense_words_loss(io, wo, c["labels"], c["cap_lens"], c["class_ids"])
    (w0 * o0 + w1 * o1).backward()
    return (l0, l1, att, img.grad, words.grad), (o0, o1, oatt, io.grad, wo.grad)


@pytest.mark.parametrize("B,T,D,H,lens", [
    (1, 18, 256, 17, [7]),                 # a single pair: CE over one cell is exactly 0
    (2, 18, 256, 17, [1, 18]),             # one-word caption: softmax over a single word is 1
    (3, 32, 64, 10, [32, 5, 17]),          # T_max = 32 (lane-per-word limit), non-CUB D / R
    (5, 9, 128, 4, [9, 9, 9, 9, 9]),       # no ragged captions, R = 16 < one TMA box
])
def test_small_and_odd_shapes(cuda_lib, B, T, D, H, lens):
    c = cases.words_case(B, T, D=D, H=H, seed=B * 100 + T, class_mode="none", min_len=1)
    c["cap_lens"] = torch.tensor(lens)
    got, ref = _run_both(c, B)
    assert abs(got[0].item() - ref[0].item()) <= 2e-5 and abs(got[1].item() - ref[1].item()) <= 2e-5
    for a, b in zip(got[2], ref[2]):
        assert tuple(a.shape) == tuple(b.shape)
        assert float((a.cpu().double() - b.detach()).abs().max()) <= 2e-6
    if B > 1:
        assert relmax(got[3].cpu(), ref[3]) <= 1e-4 and relmax(got[4].cpu(), ref[4]) <= 1e-4
    else:
        assert float(got[3].abs().max()) <= 1e-7 and float(got[4].abs().max()) <= 1e-7


@pytest.mark.parametrize("B,T,D,H,lens", [
    (3, 32, 128, 10, [32, 5, 17]),                       # T_max = 32: two captions per 64-column bin at most
    (4, 32, 128, 6, [32, 32, 32, 32]),                   # bins filled exactly (64 of 64 columns), R = 36 < 128
    (9, 20, 128, 17, [1, 2, 3, 4, 5, 8, 9, 16, 20]),     # every 4-word bucket boundary of the epilogue
    (4, 18, 512, 8, [18, 11, 6, 13]),                    # D = 512 (dU kernel NQ = 4), R = 64: no re-pitch copy
    (6, 18, 256, 20, [18, 7, 12, 5, 18, 9]),             # R = 400: four region tiles, last one 16 rows
    (7, 18, 384, 17, [18, 18, 18, 10, 18, 18, 18]),      # odd number of bins: the last tile has one live bin
])
@pytest.mark.parametrize("engine", [3, 2])
def test_fused_engine_shapes(cuda_lib, engine, B, T, D, H, lens):
    """Shapes that exercise the fused engines' packing (64-column bins, 128-column tiles), their
    per-caption word buckets and the region-tile edges, against the float64 dense oracle.
    engine 3 = half-pair operands (default), 2 = 3xTF32 with the split in the kernel."""
    assert D % 128 == 0
    default = cuda_lib.eegan_get_contraction_engine()
    assert cuda_lib.eegan_set_contraction_engine(engine) == 0
    try:
        _fused_engine_shapes(cuda_lib, B, T, D, H, lens)
    finally:
        cuda_lib.eegan_set_contraction_engine(default)


def _fused_engine_shapes(cuda_lib, B, T, D, H, lens):
    c = cases.words_case(B, T, D=D, H=H, seed=B * 100 + T + D, class_mode="none", min_len=1)
    c["cap_lens"] = torch.tensor(lens)
    got, ref = _run_both(c, B, 1.0, 0.5)
    assert abs(got[0].item() - ref[0].item()) <= 2e-5 and abs(got[1].item() - ref[1].item()) <= 2e-5
    # a map row sums to 1 over R regions: the CUB tolerance (2e-6 at R = 289) scales with the entry size 1/R
    tol_att = 2e-6 * max(1.0, 289.0 / (H * H))
    for a, b in zip(got[2], ref[2]):
        assert tuple(a.shape) == tuple(b.shape)
        assert float((a.cpu().double() - b.detach()).abs().max()) <= tol_att
        # argmax word per region: identical wherever the float64 oracle separates its top two words by more
        # than the map tolerance (an fp32 reference flips the same near-ties)
        bd = b.detach()
        top2 = bd.topk(min(2, bd.shape[1]), dim=1).values
        clear = (top2[:, 0] - top2[:, -1]) > 2 * tol_att if bd.shape[1] > 1 else torch.ones_like(top2[:, 0], dtype=torch.bool)
        assert torch.equal(a.cpu().argmax(1)[clear], bd.argmax(1)[clear])
    assert relmax(got[3].cpu(), ref[3]) <= 1e-4 and relmax(got[4].cpu(), ref[4]) <= 1e-4
    for i, n in enumerate(lens):  # padded words receive exactly zero gradient
        assert float(got[4][i, :, n:].abs().max()) == 0.0 if n < T else True


@pytest.mark.parametrize("img_amp,word_amp,loss_amp", [(1e-3, 1e-3, 1.0), (20.0, 0.05, 1.0), (1.0, 1.0, 1e-6), (1.0, 1.0, 3e4),
                                                         (1e-2, 40.0, 1e-3)])
def test_half_pair_engine_operand_scales(cuda_lib, img_amp, word_amp, loss_amp):
    """The half-pair engine stores every GEMM operand as fp16 hi/lo of x * 2^e: inputs and upstream gradients far from
    unit magnitude must keep the fp32-class accuracy (the scales come from device-side maxima / bounds, pair_grid_h.cu)."""
    assert cuda_lib.eegan_get_contraction_engine() == 3
    B, T = 9, 18
    c = cases.words_case(B, T, seed=91, class_mode="cub")
    c["img"] = c["img"] * img_amp
    c["words"] = c["words"] * word_amp
    got, ref = _run_both(c, B, loss_amp, 0.5 * loss_amp)
    assert abs(got[0].item() - ref[0].item()) <= 2e-5 * max(1.0, abs(ref[0].item()))
    assert abs(got[1].item() - ref[1].item()) <= 2e-5 * max(1.0, abs(ref[1].item()))
    assert relmax(got[3].cpu(), ref[3]) <= 1e-4 and relmax(got[4].cpu(), ref[4]) <= 1e-4
    for a, b in zip(got[2], ref[2]):
        assert float((a.cpu().double() - b.detach()).abs().max()) <= 2.5e-6


def test_nan_input_propagates_like_the_reference(cuda_lib):
    """A NaN in a live input element makes both losses NaN in the reference (it flows through bmm / softmax / CE); the
    half-pair operand conversion must not launder it into a finite value."""
    import eegan_b200 as E
    c = cases.words_case(5, 18, seed=17, class_mode="none")
    c["img"][2, 7, 3, 4] = float("nan")
    l0, l1, _ = E.words_loss(c["img"].cuda(), c["words"].cuda(), c["labels"].cuda(), c["cap_lens"].cuda(), None, 5)
    o0, o1, _ = O.port_words_loss(c["img"], c["words"], c["labels"], c["cap_lens"], None, 5)
    assert torch.isnan(o0) and torch.isnan(o1)
    assert torch.isnan(l0) and torch.isnan(l1)


def test_backward_twice_on_one_forward(cuda_lib):
    """retain_graph: the backward leaves the forward stash intact (dU goes to its own buffer)."""
    import eegan_b200 as E
    c = cases.words_case(6, 18, seed=77)
    img = c["img"].cuda().requires_grad_()
    words = c["words"].cuda().requires_grad_()
    l0, l1, _ = E.words_loss(img, words, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], 6)
    (l0 + l1).backward(retain_graph=True)
    g1 = (img.grad.clone(), words.grad.clone())
    img.grad = None
    words.grad = None
    (l0 + l1).backward()
    assert torch.equal(img.grad, g1[0]) and torch.equal(words.grad, g1[1])


def test_all_same_class_masks_every_off_diagonal(cuda_lib):
    """class_ids all equal: every off-diagonal cell is -inf (DAMSM_losses.py:282-285,331-333), each
    CE row has one finite logit, both losses are exactly 0 and no gradient flows."""
    import eegan_b200 as E
    B = 6
    c = cases.words_case(B, 12, seed=9)
    c["class_ids"] = torch.full((B,), 7)
    img = c["img"].cuda().requires_grad_()
    words = c["words"].cuda().requires_grad_()
    l0, l1, _ = E.words_loss(img, words, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
    (l0 + l1).backward()
    assert l0.item() == 0.0 and l1.item() == 0.0
    assert float(img.grad.abs().max()) == 0.0 and float(words.grad.abs().max()) == 0.0
    sim, _ = E.words_similarity(img.detach(), words.detach(), c["cap_lens"].cuda(), c["class_ids"], B)
    off = ~torch.eye(B, dtype=torch.bool, device="cuda")
    assert torch.isinf(sim[off]).all() and torch.isfinite(sim.diagonal()).all()


def test_batch_larger_than_one_tile_everywhere(cuda_lib):
    """B = 130: > 128 images, sum(cap_lens) spans many 128-row tiles, GEMM5 splits images unevenly.
    Checked against the fp32 loop port (the float64 dense oracle would need > 10 GB here)."""
    B, T = 130, 18
    c = cases.words_case(B, T, seed=4, class_mode="cub")
    got, ref = _run_both(c, B, 1.0, 2.0, use_port=True)
    assert abs(got[0].item() - ref[0].item()) <= 5e-5 * abs(ref[0].item())
    assert abs(got[1].item() - ref[1].item()) <= 5e-5 * abs(ref[1].item())
    assert relmax(got[3].cpu(), ref[3]) <= 1e-4 and relmax(got[4].cpu(), ref[4]) <= 1e-4
    for a, b in zip(got[2][:8], ref[2][:8]):
        assert float((a.cpu() - b.detach()).abs().max()) <= 1e-6


def test_sharded_block_equals_columns_of_full_grid(cuda_lib):
    """The caption-row-sharded building block on ONE GPU: a rank's [B, b] column block with
    diag_offset must equal the same columns of the full grid, and its att maps the matching rows."""
    from eegan_b200.damsm_losses import pair_grid
    B, b = 12, 4
    c = cases.words_case(B, 18, seed=2)
    img, words, lens = c["img"].cuda(), c["words"].cuda(), c["cap_lens"].cuda()
    full, att_full = pair_grid(img, words, lens)
    for rank in range(B // b):
        sl = slice(rank * b, (rank + 1) * b)
        blk, att = pair_grid(img, words[sl], lens[sl], diag_offset=rank * b)
        assert float((blk - full[:, sl]).abs().max()) <= 2e-6
        assert float((att - att_full[sl]).abs().max()) <= 1e-7


@pytest.mark.parametrize("B,b,T", [(12, 4, 18), (384, 48, 18), (96, 32, 20)])
def test_sharded_block_equals_rows_of_full_grid(cuda_lib, B, b, T):
    """The image-row partition (the default of eegan_b200.sharded) on ONE GPU: a rank's block = its b images against ALL B captions
    with a negative diag_offset.  Forward: the rows of the full grid, attention maps only for the rank's own captions (zeros
    elsewhere).  Backward with the rank's rows of dm: d_img of the own images is complete, the partial d_words of the blocks add up
    to the full-batch d_words (what the reduce-scatter does).  (384, 48) is the block of the 8-GPU CUB run: B_img = 48, B_cap = 384."""
    from eegan_b200.damsm_losses import pair_grid
    c = cases.words_case(B, T, seed=B + b)
    img, words, lens = c["img"].cuda().requires_grad_(), c["words"].cuda().requires_grad_(), c["cap_lens"].cuda()
    full, att_full = pair_grid(img, words, lens)
    dm = torch.randn(B, B, generator=cases._gen(3)).cuda() * 0.1
    (full * dm).sum().backward()
    d_img_full, d_words_full = img.grad.clone(), words.grad.clone()
    d_words_sum = torch.zeros_like(d_words_full)
    ranks = range(B // b) if B <= 96 else (0, 3, B // b - 1)  # the large case: three of the eight blocks (no d_words sum check)
    for rank in ranks:
        sl = slice(rank * b, (rank + 1) * b)
        ib, wb = img.detach()[sl].clone().requires_grad_(), words.detach().clone().requires_grad_()
        blk, att = pair_grid(ib, wb, lens, diag_offset=-rank * b)
        assert tuple(blk.shape) == (b, B) and float((blk - full[sl]).abs().max()) <= 2e-6
        assert float((att[sl] - att_full[sl]).abs().max()) <= 1e-7
        rest = torch.ones(B, dtype=torch.bool)
        rest[sl] = False
        assert float(att[rest.cuda()].abs().max()) == 0.0
        (blk * dm[sl]).sum().backward()
        assert relmax(ib.grad.cpu(), d_img_full[sl].cpu()) <= 2e-5
        d_words_sum += wb.grad
    if B <= 96:
        assert relmax(d_words_sum.cpu(), d_words_full.cpu()) <= 2e-5


@pytest.mark.parametrize("B,Rv,D", [(10, 100, 256), (3, 7, 64), (1, 1, 256)])
def test_r_precision_matches_reference_arithmetic(cuda_lib, B, Rv, D):
    """eegan_b200.r_precision vs the restated loop body of Tester.cal_sim_one_by_one (test.py:323-330):
    scores within fp32 noise, argmax / hit bit-exact (ties included: lowest index)."""
    import eegan_b200 as E
    g = cases._gen(B * 1000 + Rv)
    cnn = torch.randn(B, D, generator=g)
    rnn = torch.randn(B, Rv, D, generator=g)
    if Rv > 1:
        rnn[0, 0] = cnn[0] * 2.0                 # a clear hit
    if Rv > 3:
        rnn[1 % B, 3] = rnn[1 % B, 2]            # an exact tie between two candidates
    if B > 2:
        rnn[2, :, :] = 0.0                       # all-zero candidates: the clamp (min=1e-8) path, scores 0
    hits, best, scores = E.r_precision(cnn.cuda(), rnn.cuda(), return_scores=True)
    oh, ob, osc = O.port_rprecision(cnn, rnn)
    assert float((scores.cpu() - osc).abs().max()) <= 2e-6
    # argmax: exact wherever the reference separates its top two by more than the score tolerance
    top2 = osc.topk(min(2, Rv), dim=1).values
    clear = (top2[:, 0] - top2[:, -1]) > 4e-6 if Rv > 1 else torch.ones(B, dtype=torch.bool)
    assert torch.equal(best.cpu().long()[clear], ob[clear]) and torch.equal(hits.cpu()[clear], oh[clear])
    # exact ties resolve to the lowest index, as torch.argmax does
    if Rv > 3:
        s1 = scores[1 % B].cpu()
        assert s1[2] == s1[3]
    if B > 2:
        assert int(best[2]) == 0 and bool(hits[2])


def test_phased_backward_equals_one_call(cuda_lib):
    """eegan_damsm_pair_bwd_phased: phase 1 (d_img) then phase 2 (d_words) on the same stash == eegan_damsm_pair_bwd."""
    from eegan_b200 import _lib
    L = cuda_lib
    B, T, D, R = 20, 14, 256, 289
    c = cases.words_case(B, T, seed=8)
    img = c["img"].cuda().reshape(B, D, R).contiguous()
    words = c["words"].cuda().contiguous()
    lens = c["cap_lens"].cuda().to(torch.int32)
    need = L.eegan_damsm_pair_workspace_bytes(B, B, D, R, T)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    m, att = torch.empty(B, B, device="cuda"), torch.empty(B, T, R, device="cuda")
    p, st = _lib.ptr, _lib.stream_ptr()
    _lib.check(L.eegan_damsm_pair_fwd(p(img), p(words), p(lens), B, B, D, R, T, 5.0, 5.0, p(m), p(att), 0, p(ws), need, st))
    dm = torch.randn(B, B, generator=cases._gen(9)).cuda() * 0.01
    di0, dw0 = torch.empty_like(img), torch.empty_like(words)
    _lib.check(L.eegan_damsm_pair_bwd(p(img), p(words), p(lens), B, B, D, R, T, 5.0, 5.0, p(dm), p(di0), p(dw0), p(ws), need, st))
    di1, dw1 = torch.zeros_like(img), torch.zeros_like(words)
    _lib.check(L.eegan_damsm_pair_bwd_phased(p(img), p(words), p(lens), B, B, D, R, T, 5.0, 5.0, p(dm), p(di1), None, 1, p(ws), need, st))
    assert float(dw1.abs().max()) == 0.0
    _lib.check(L.eegan_damsm_pair_bwd_phased(p(img), p(words), p(lens), B, B, D, R, T, 5.0, 5.0, p(dm), None, p(dw1), 2, p(ws), need, st))
    assert torch.equal(di0, di1) and torch.equal(dw0, dw1)
    assert L.eegan_damsm_pair_bwd_phased(p(img), p(words), p(lens), B, B, D, R, T, 5.0, 5.0, p(dm), p(di1), p(dw1), 7, p(ws), need, st) != 0
