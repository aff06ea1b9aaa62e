"""acx_gemm with the A operand generated in the kernel from uint8 observations (a_patch_u8: the conv1 patch matrix is never
materialised) against the same GEMM on the materialised patch matrix: same bf16 operand values and the same accumulation
order in the forward product, so that result must be bit-identical; the MN-major products (weight gradient, input factor)
split K differently since the materialised operand runs 128-deep k-blocks, so they agree to fp32 summation order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _patches(obs):
    """uint8 [S,84,84,4] -> bf16 [S*400, 256], rows (r, oy, ox), columns (kh, kw, c)  (extract_image_patches order)."""
    s = obs.shape[0]
    x = obs.float().permute(0, 3, 1, 2)                                   # [S,4,84,84]
    cols = torch.nn.functional.unfold(x, kernel_size=8, stride=4)          # [S, 4*64, 400], feature order (c, kh, kw)
    cols = cols.view(s, 4, 8, 8, 400).permute(0, 4, 2, 3, 1)               # [S, 400, kh, kw, c]
    return cols.reshape(s * 400, 256).to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("samples", [3, 37])
def test_patch_operand_forward_wgrad_syrk_bit_identical(samples):
    from actorcritic_b200 import ops, _lib
    gen = torch.Generator(device="cuda").manual_seed(samples)
    obs = torch.randint(0, 256, (samples, 84, 84, 4), dtype=torch.uint8, device="cuda", generator=gen)
    p1 = _patches(obs)
    rows = samples * 400
    # forward: [rows, 256] x W^T [32, 256] (3 planes), bias + ReLU, 3 output planes
    w = ops.split_planes(torch.randn((32, 256), device="cuda", generator=gen) * 0.05, 3)
    bias = torch.randn(32, device="cuda", generator=gen)
    pairs = [(0, 0), (0, 1), (0, 2)]
    _, want = ops.gemm([p1], w, rows, 32, 256, pairs=pairs, alpha=1 / 255.0, bias=bias, relu=True, out_planes=3, want_f32=False)
    _, got = ops.gemm(None, w, rows, 32, 256, pairs=pairs, alpha=1 / 255.0, bias=bias, relu=True, out_planes=3, want_f32=False,
                      a_patch_obs=obs)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    # wgrad: P1^T [256, rows] x g [rows, 32] (3 planes)
    g = ops.split_planes(torch.randn((rows, 32), device="cuda", generator=gen) * 1e-3, 3)
    want, _ = ops.gemm([p1], g, 256, 32, rows, trans=True, pairs=pairs, alpha=1 / 255.0)
    got, _ = ops.gemm(None, g, 256, 32, rows, trans=True, pairs=pairs, alpha=1 / 255.0, a_patch_obs=obs)
    assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())
    # input factor: P1^T P1 (symmetric, panel mode)
    want, _ = ops.gemm([p1], [p1], 256, 256, rows, trans=True, symmetric=True, pairs=[(0, 0)], alpha=1.0 / rows)
    got, _ = ops.gemm(None, None, 256, 256, rows, trans=True, symmetric=True, alpha=1.0 / rows, a_patch_obs=obs)
    assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())
    assert _lib.load().acx_debug_tc_error() == 0
    # and against plain arithmetic (exact integers up to fp32 accumulation)
    ref = (p1.double().t() @ p1.double() / rows).cpu().numpy()
    assert np.abs(got.double().cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
