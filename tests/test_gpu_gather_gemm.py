"""acx_gemm with operands read in place from NHWC tensors (acx_gather_t, include/acx.h) against the same products on
materialised patch matrices: the factor statistic P^T P kfac builds from extract_image_patches
(envs/atari/model.py:227-237) and the weight gradient P^T g of nn.conv2d (nn.py:88-110, objectives.py:79) for the three
convolutions of the Nature-CNN (8x8/4 on [84,84,4], 4x4/2 on [20,20,32], 3x3/1 on [9,9,64]).  The k-blocks enumerate the
locations in another order than the rows of a patch matrix, so the results agree to fp32 rounding, not bit for bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _patch_matrix(x_nhwc, k, s):
    """[S,H,W,C] (double) -> [S*T, k*k*C], rows (sample, oy, ox), columns (kh, kw, c)."""
    S, H, W, C = x_nhwc.shape
    cols = torch.nn.functional.unfold(x_nhwc.permute(0, 3, 1, 2), kernel_size=k, stride=s)   # [S, C*k*k, T]: (c, kh, kw)
    T = cols.shape[-1]
    return cols.view(S, C, k, k, T).permute(0, 4, 2, 3, 1).reshape(S * T, k * k * C)


def _gathers(ops, planes, geom, samples):
    """The in-place views of csrc/learner.cu:setup_gather for one of the three geometries."""
    if geom == "conv2":      # act1 [S,20,20,32]: chunk (kh, h) -> row 2 oy + kh, pixel 2 (ox + h)
        px = 32 * 2
        chunks = [(0, h, kh, 0) for kh in range(4) for h in range(2)]
        return ops.gather_view(planes, (64, 10, 4, 9, samples), (2 * px, 20 * px, 40 * px, 400 * px), 9, 9, samples, chunks)
    if geom == "conv3":      # act2 [S,9,9,64]: chunk (kh, kw) -> pixel (oy + kh, ox + kw)
        px = 64 * 2
        chunks = [(0, kw, 0, kh) for kh in range(3) for kw in range(3)]
        return ops.gather_view(planes, (64, 9, 1, 9, samples), (px, px, 9 * px, 81 * px), 7, 7, samples, chunks)
    prow = 84 * 8 * 2        # conv1 on the row-pair copy [S,42,84,2,4]: chunk j -> pair-row 2 oy + j, pixel 4 ox
    return ops.gather_view(planes, (64, 20, 4, 20, samples), (64, prow, 2 * prow, 42 * prow), 20, 20, samples,
                           [(0, 0, j, 0) for j in range(4)])


def _grad_gather(ops, planes, c, hw_out, samples):
    px = c * 2
    return ops.gather_view(planes, (c, hw_out, 1, hw_out, samples), (px, px, hw_out * px, hw_out * hw_out * px), hw_out, hw_out,
                           samples, [(64 * j, 0, 0, 0) for j in range((c + 63) // 64)])


GEOMS = {"conv2": (20, 32, 4, 2, 9, 64), "conv3": (9, 64, 3, 1, 7, 32)}


@pytest.mark.parametrize("samples", [5, 64, 200])
@pytest.mark.parametrize("geom", ["conv2", "conv3"])
def test_gathered_factor_and_weight_gradient_match_the_patch_matrix(geom, samples):
    from actorcritic_b200 import _lib, ops
    hw_in, c_in, k, s, hw_out, c_out = GEOMS[geom]
    gen = torch.Generator(device="cuda").manual_seed(samples + hw_in)
    x = torch.rand((samples * hw_in * hw_in, c_in), device="cuda", generator=gen)
    xp = ops.split_planes(x, 2)
    xsum = sum(p.double() for p in xp).view(samples, hw_in, hw_in, c_in)
    P = _patch_matrix(xsum, k, s)
    K, rows = k * k * c_in, samples * hw_out * hw_out
    ga = _gathers(ops, xp, geom, samples)
    pairs = ops.PAIRS[3]
    # input factor P^T P / rows
    got, _ = ops.gemm(xp, xp, K, K, rows, trans=True, pairs=pairs, symmetric=True, alpha=1.0 / rows, a_gather=ga)
    hi, lo = (_patch_matrix(p.double().view(samples, hw_in, hw_in, c_in), k, s) for p in xp)
    want = (hi.t() @ hi + hi.t() @ lo + lo.t() @ hi) / rows     # the three plane pairs the kernel accumulates
    assert float((got.double() - want).abs().max()) <= 2e-5 * float(want.abs().max())
    assert float((got.double() - P.t() @ P / rows).abs().max()) <= 1e-4 * float(want.abs().max())
    # weight gradient P^T g
    g = torch.randn((rows, c_out), device="cuda", generator=gen) * 1e-2
    gp = ops.split_planes(g, 2)
    gb = _grad_gather(ops, gp, c_out, hw_out, samples)
    got, _ = ops.gemm(xp, gp, K, c_out, rows, trans=True, pairs=pairs, a_gather=ga, b_gather=gb)
    ghi, glo = (p.double()[:, :c_out] for p in gp)
    want = hi.t() @ ghi + hi.t() @ glo + lo.t() @ ghi
    assert float((got.double() - want).abs().max()) <= 2e-5 * float(want.abs().max()) + 1e-9
    assert _lib.load().acx_debug_tc_error() == 0


@pytest.mark.parametrize("samples", [3, 40])
def test_conv1_from_the_row_pair_copy(samples):
    from actorcritic_b200 import _lib, ops
    gen = torch.Generator(device="cuda").manual_seed(samples)
    obs = torch.randint(0, 256, (samples, 84, 84, 4), dtype=torch.uint8, device="cuda", generator=gen)
    pairs_copy = ops.obs_pairs(obs)
    want_copy = obs.view(samples, 42, 2, 84, 4).permute(0, 1, 3, 2, 4).to(torch.bfloat16)
    assert torch.equal(pairs_copy, want_copy.contiguous())
    P = _patch_matrix(obs.double(), 8, 4)                       # [S*400, 256], columns (kh, kw, c): exact integers
    rows = samples * 400
    ga = _gathers(ops, [pairs_copy], "conv1", samples)
    # input factor: panel mode, rows and columns stored back in the (kh, kw, c) order (perm_m, perm_n)
    got, _ = ops.gemm([pairs_copy], [pairs_copy], 256, 256, rows, trans=True, pairs=[(0, 0)], symmetric=True, alpha=1.0 / rows,
                      a_gather=ga, perm_m=1, perm_n=1)
    want = P.t() @ P / rows
    assert float((got.double() - want).abs().max()) <= 1e-6 * float(want.abs().max())
    # weight gradient: rows stored back in the (kh, kw, c) order
    g = torch.randn((rows, 32), device="cuda", generator=gen) * 1e-3
    gp = ops.split_planes(g, 2)
    gb = _grad_gather(ops, gp, 32, 20, samples)
    got, _ = ops.gemm([pairs_copy], gp, 256, 32, rows, trans=True, pairs=[(0, 0), (0, 1)], alpha=1 / 255.0, a_gather=ga, b_gather=gb,
                      perm_m=1)
    want = P.t() @ sum(p.double()[:, :32] for p in gp) / 255.0
    assert float((got.double() - want).abs().max()) <= 2e-6 * float(want.abs().max())
    assert _lib.load().acx_debug_tc_error() == 0


@pytest.mark.parametrize("samples", [2, 33])
def test_conv1_forward_on_the_row_pair_copy(samples):
    """acx_conv1_pairs_forward (envs/atari/model.py:173-179: 8x8 / 4 convolution of obs / 255, bias, ReLU) against conv2d in fp64."""
    from actorcritic_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(samples)
    obs = torch.randint(0, 256, (samples, 84, 84, 4), dtype=torch.uint8, device="cuda", generator=gen)
    w = torch.randn((256, 32), device="cuda", generator=gen) * 0.05            # rows (kh, kw, c): the HWIO kernel flattened
    bias = torch.randn(32, device="cuda", generator=gen) * 0.1
    outs = ops.conv1_pairs_forward(ops.obs_pairs(obs), w, bias, w_planes=3, pairs=[(0, 0), (0, 1), (0, 2)])
    got = sum(p.double() for p in outs)
    wk = w.double().view(8, 8, 4, 32).permute(3, 2, 0, 1)                     # OIHW
    want = torch.nn.functional.conv2d(obs.double().permute(0, 3, 1, 2) / 255.0, wk, bias.double(), stride=4).clamp_min(0)
    want = want.permute(0, 2, 3, 1).reshape(samples * 400, 32)
    assert float((got - want).abs().max()) <= 3e-5 * float(want.abs().max())   # two bf16 output planes: 2^-16
