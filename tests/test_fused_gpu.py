"""GPU tests of the tensor-core (tcgen05) path.

Numerics contract of the fused path: operands of every tensor-core contraction (activations and
BatchNorm-folded weights) are rounded to bf16 (RNE), products are exact and accumulate in fp32.
`mixed_reference` below restates exactly that in torch; the kernels must agree with it to
rtol 1e-3 (differences: fp32 summation order, and rare one-ulp bf16 flips of intermediate
activations).  Against the plain fp32 reference the same outputs agree to ~1e-2 of the tensor's
max magnitude (bf16 has 8 mantissa bits); both tolerances are asserted."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (1000, 128, 100), (4096, 256, 512), (130, 16, 64), (77, 256, 259),
                                   (300, 128, 128)])
@pytest.mark.parametrize("out", ["f32", "bf16", "cm"])
def test_linear_bf16(pkg, cuda, m, n, k, out):
    import importlib
    tcm = importlib.import_module("3dvlp_b200.tensorcore")
    g = torch.Generator().manual_seed(m * 7 + n)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda)
    bias = torch.randn(n, generator=g).to(cuda)
    x = torch.randn(m, k, generator=g).to(cuda)
    pw = tcm.PackedWeight(w)
    xp = torch.zeros(m, pw.kpad, device=cuda, dtype=torch.bfloat16)
    xp[:, :k] = x.to(torch.bfloat16)
    rpb = m // 2 if (out == "cm" and m % 2 == 0) else m
    y = tcm.linear(xp, pw, bias, relu=True, out=out, rows_per_batch=rpb)
    assert tcm.poll_device_error(cuda) == 0
    want = torch.relu(bf16r(x).double() @ bf16r(w).double().t() + bias.double()).float()
    if out == "cm":
        y = y.permute(0, 2, 1).reshape(m, n)
    if out == "bf16":
        torch.testing.assert_close(y.float(), want, rtol=1e-2, atol=1e-2)
    else:
        torch.testing.assert_close(y, want, rtol=1e-3, atol=1e-4)


def test_pack_rows(pkg, cuda):
    import importlib
    tcm = importlib.import_module("3dvlp_b200.tensorcore")
    f = torch.randn(3, 131, 777, device=cuda)
    p = tcm.pack_rows(f)
    assert p.shape == (3 * 777, 192)
    want = torch.zeros(3, 777, 192, device=cuda)
    want[:, :, :131] = f.transpose(1, 2)
    assert torch.equal(p.float().view(3, 777, 192), bf16r(want))
