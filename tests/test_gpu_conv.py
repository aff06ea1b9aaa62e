"""acx_conv (implicit-GEMM forward, gather-form input gradient) against a plain fp64 torch convolution of the same
bf16-plane operands.  The operands are the exact sums of their planes, so the only error left is the fp32 accumulation of
the tensor-core kernel: tolerance 1e-5 norm-relative, 2e-3 guarded element-wise (6 plane pairs drop the 2^-24-relative cross terms; outputs
are cancelling sums; 3 output planes carry the fp32 result exactly)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CONV2 = (20, 32, 4, 2, 9, 64)     # envs/atari/model.py:186-191
CONV3_32 = (9, 64, 3, 1, 7, 32)   # envs/atari/model.py:193-199 with conv3_num_filters=32 (a2c_acktr.py:52)
CONV3_64 = (9, 64, 3, 1, 7, 64)


def _ops():
    from actorcritic_b200 import ops
    return ops


def _planes_of(x, shape):
    """fp32 tensor -> (list of 3 bf16 planes shaped `shape`, their exact fp64 sum on the CPU)."""
    ops = _ops()
    flat = x.reshape(-1, shape[-1]).contiguous()
    planes = ops.split_planes(flat, 3)
    assert planes[0].shape[1] == shape[-1]
    total = sum(p.double() for p in planes).cpu().reshape(shape)
    return [p.reshape(shape) for p in planes], total


def _errs(got, want):
    got, want = got.double().cpu().numpy(), want.numpy()
    rel = np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30)
    floor = 1e-3 * np.abs(want).max()
    elem = (np.abs(got - want) / np.maximum(np.abs(want), floor)).max()
    return rel, elem


@pytest.mark.parametrize("geom,samples", [(CONV2, 5), (CONV3_32, 5), (CONV3_64, 4), (CONV2, 150), (CONV3_32, 301)])
def test_conv_forward_matches_fp64(geom, samples):
    ops = _ops()
    hw_in, c_in, k, s, hw_out, c_out = geom
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand((samples, hw_in, hw_in, c_in), device="cuda", generator=gen) * (torch.rand((samples, hw_in, hw_in, c_in), device="cuda", generator=gen) > 0.3)
    w = torch.randn((k * k * c_in, c_out), device="cuda", generator=gen) * 0.05
    bias = torch.randn((c_out,), device="cuda", generator=gen) * 0.1
    xp, xs = _planes_of(x, (samples, hw_in, hw_in, c_in))
    wtp, wts = _planes_of(w.t().contiguous(), (c_out, k * k * c_in))
    outs = ops.conv(xp, wtp, geom, samples, bias=bias, relu=True)
    torch.cuda.synchronize()
    from actorcritic_b200 import _lib
    assert _lib.load().acx_debug_tc_error() == 0
    w_oihw = wts.reshape(c_out, k, k, c_in).permute(0, 3, 1, 2)
    want = F.relu(F.conv2d(xs.permute(0, 3, 1, 2), w_oihw, bias.double().cpu(), stride=s)).permute(0, 2, 3, 1)
    got = sum(p.double() for p in outs)
    rel, elem = _errs(got, want)
    assert rel <= 1e-5 and elem <= 2e-3, (rel, elem)


@pytest.mark.parametrize("geom,samples,mask_samples", [(CONV3_32, 6, 3), (CONV3_64, 4, 4), (CONV2, 4, 2), (CONV2, 160, 80),
                                                        (CONV3_32, 300, 150)])
def test_conv_dgrad_matches_fp64(geom, samples, mask_samples):
    ops = _ops()
    hw_in, c_in, k, s, hw_out, c_out = geom
    gen = torch.Generator(device="cuda").manual_seed(5)
    g = torch.randn((samples, hw_out, hw_out, c_out), device="cuda", generator=gen)
    w = torch.randn((k * k * c_in, c_out), device="cuda", generator=gen) * 0.05
    act = torch.randn((mask_samples, hw_in, hw_in, c_in), device="cuda", generator=gen).clamp_min(0).to(torch.bfloat16)
    gp, gs = _planes_of(g, (samples, hw_out, hw_out, c_out))
    wd = ops.conv_dgrad_weights(w, geom)
    outs = ops.conv(gp, wd, geom, samples, dgrad=True, mask=act, mask_samples=mask_samples)
    torch.cuda.synchronize()
    from actorcritic_b200 import _lib
    assert _lib.load().acx_debug_tc_error() == 0
    # the kernel consumes the bf16 planes of the rearranged weights: rebuild W from them so that the reference sees the
    # same operand bits
    m = k // s
    wsum = sum(p.double() for p in wd).cpu()[:, :m * m * c_out].reshape(s, s, c_in, m, m, c_out)   # (py,px,ci,i,j,co)
    w_hwio = torch.zeros((k, k, c_in, c_out), dtype=torch.float64)
    for py in range(s):
        for px in range(s):
            for i in range(m):
                for j in range(m):
                    w_hwio[s * i + py, s * j + px] = wsum[py, px, :, i, j, :]
    w_oihw = w_hwio.permute(3, 2, 0, 1)
    dx = F.conv_transpose2d(gs.permute(0, 3, 1, 2), w_oihw, stride=s).permute(0, 2, 3, 1)
    assert tuple(dx.shape) == (samples, hw_in, hw_in, c_in)
    keep = (act.double().cpu() > 0).repeat(samples // mask_samples, 1, 1, 1)
    want = dx * keep
    got = sum(p.double() for p in outs)
    rel, elem = _errs(got, want)
    assert rel <= 1e-5 and elem <= 2e-3, (rel, elem)


def test_conv_dgrad_weights_layout():
    ops = _ops()
    hw_in, c_in, k, s, hw_out, c_out = CONV2
    w = torch.arange(k * k * c_in * c_out, device="cuda", dtype=torch.float32).reshape(k * k * c_in, c_out) % 251
    planes = ops.conv_dgrad_weights(w, CONV2)
    got = planes[0].float().cpu().numpy()       # values < 256 are exact in the hi plane
    w4 = w.cpu().numpy().reshape(k, k, c_in, c_out)
    m = k // s
    for (py, px, ci, i, j, co) in [(0, 0, 0, 0, 0, 0), (1, 0, 5, 0, 1, 7), (1, 1, 31, 1, 1, 63), (0, 1, 17, 1, 0, 40)]:
        assert got[(py * s + px) * c_in + ci, (i * m + j) * c_out + co] == w4[s * i + py, s * j + px, ci, co]


def test_conv_rejects_unsupported_geometry():
    ops = _ops()
    from actorcritic_b200 import _lib
    x = [torch.zeros((2, 84, 84, 4), dtype=torch.bfloat16, device="cuda")]
    w = [torch.zeros((32, 256), dtype=torch.bfloat16, device="cuda")]
    with pytest.raises(_lib.AcxError):
        ops.conv(x, w, (84, 4, 8, 4, 20, 32), 2)
