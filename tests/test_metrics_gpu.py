"""Loss / softmax / top-k counter kernels vs torch and the reference's accuracy()."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ce_loss_and_grad(cuda):
    from robustart_b200 import ops
    torch.manual_seed(0)
    z = (torch.randn(257, 1000, device=cuda) * 5)
    y = torch.randint(0, 1000, (257,), device=cuda)
    loss, d = ops.ce_loss_grad(z, y)
    zr = z.clone().requires_grad_(True)
    lr = torch.nn.functional.cross_entropy(zr, y, reduction="none")
    lr.sum().backward()
    assert (loss - lr.detach()).abs().max().item() < 1e-4
    assert (d - zr.grad).abs().max().item() < 1e-6
    sm = ops.softmax(z)
    assert (sm - torch.softmax(z, 1)).abs().max().item() < 1e-6


def test_topk_counters_bit_exact(cuda):
    from robustart_b200 import ops
    from oracle import metrics as OM
    torch.manual_seed(1)
    n = 1000
    z = torch.randn(n, 1000, device=cuda)
    y = torch.randint(0, 1000, (n,), device=cuda)
    y[:300] = z[:300].argmax(1)                      # force top-1 hits
    idx5 = z[300:500].topk(5, 1).indices[:, 4]       # force rank-5 hits
    y[300:500] = idx5
    counters = torch.zeros(3, dtype=torch.int64, device=cuda)
    pred = torch.empty(n, dtype=torch.int64, device=cuda)
    ops.topk_count_(counters, z, y, pred)
    ops.topk_count_(counters, z[:10].contiguous(), y[:10].contiguous())   # accumulates
    h1, h5 = OM.topk_hits(z.cpu(), y.cpu())
    h1b, h5b = OM.topk_hits(z[:10].cpu(), y[:10].cpu())
    assert counters.tolist() == [h1 + h1b, h5 + h5b, n + 10]
    assert torch.equal(pred, z.argmax(1))
    acc1, acc5 = OM.accuracy(z.cpu(), y.cpu(), (1, 5))
    assert abs(acc1.item() - 100.0 * h1 / n) < 1e-4 and abs(acc5.item() - 100.0 * h5 / n) < 1e-4


def test_topk_ties_lowest_index(cuda):
    from robustart_b200 import ops
    z = torch.zeros(2, 1000, device=cuda)
    y = torch.tensor([0, 7], device=cuda)
    c = torch.zeros(3, dtype=torch.int64, device=cuda)
    p = torch.empty(2, dtype=torch.int64, device=cuda)
    ops.topk_count_(c, z, y, p)
    assert p.tolist() == [0, 0]
    assert c.tolist() == [1, 1, 2]   # label 0: rank 0 ; label 7: rank 7 -> miss


def test_layout_and_normalize(cuda):
    from robustart_b200 import ops
    import numpy as np
    rs = np.random.RandomState(0)
    img = torch.from_numpy(rs.randint(0, 256, size=(5, 224, 224, 3), dtype=np.uint8)).to(cuda)
    out = ops.u8nhwc_to_f32nchw(img)
    m = torch.tensor(ops.IMAGENET_MEAN, device=cuda).view(1, 3, 1, 1)
    s = torch.tensor(ops.IMAGENET_STD, device=cuda).view(1, 3, 1, 1)
    # ToTensor + Normalize run on the CPU in the reference's loader (true division by 255; torch's CUDA
    # kernel multiplies by the reciprocal instead), so the bit-exact reference is the CPU computation.
    ref = ((img.cpu().permute(0, 3, 1, 2).float() / 255) - m.cpu()) / s.cpu()
    assert torch.equal(out.cpu(), ref)
    for shape in [(2, 48, 48, 3), (1, 52, 44, 3), (3, 4, 4, 3), (2, 256, 256, 3)]:     # full CTAs, partial last CTA, a single group
        im = torch.from_numpy(rs.randint(0, 256, size=shape, dtype=np.uint8)).to(cuda)
        want = ((im.cpu().permute(0, 3, 1, 2).float() / 255) - m.cpu()) / s.cpu()
        assert torch.equal(ops.u8nhwc_to_f32nchw(im).cpu(), want)
    inv = ops.normalize(out, "inv")
    assert torch.equal(inv, out * s + m)
    assert torch.equal(ops.normalize(inv, "normal"), (inv - m) / s)
