"""ResNet forward on the tcgen05 kernels vs golden logits from the REFERENCE's own model classes
(tests/golden/make_golden_models.py; CPU fp32, weights = nets.random_state_dict(spec, 0)).
Tolerance (BASELINE north_star): logits within 1e-3 absolute; argmax / top-5 membership bit-exact."""
import os

import numpy as np
import pytest
import torch

from util import synth_images

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "resnet_logits.npz"))


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_logits_match_reference(cuda, arch):
    from robustart_b200 import nets
    model = nets.build_model(arch, device=cuda, seed=0)
    images = torch.from_numpy(synth_images(4, seed=7)).to(cuda)
    logits = model(images)
    torch.cuda.synchronize()
    got = logits.cpu().numpy()
    want = GOLD[arch]
    err = np.abs(got - want).max()
    assert err < 1e-3, (arch, err)
    assert err / np.abs(want).max() < 2e-3, (arch, err / np.abs(want).max())
    assert (got.argmax(1) == want.argmax(1)).all()
    for g, w in zip(got, want):
        assert set(np.argsort(-g)[:5]) == set(np.argsort(-w)[:5])
    # the float (attack-path) entry gives the same logits as the uint8 entry
    x01 = images.permute(0, 3, 1, 2).float().div(255).contiguous()
    got2 = model(x01).cpu().numpy()
    assert np.abs(got2 - got).max() < 1e-4
    # plain bf16 (passes = 1) is NOT within tolerance of fp32 -- that is why the default is split-bf16
    fast = nets.build_model(arch, device=cuda, seed=0, passes=1)(images).cpu().numpy()
    assert np.abs(fast - want).max() < 0.1


def test_graph_replay_and_counters(cuda):
    from robustart_b200 import nets, ops
    model = nets.build_model("resnet18", device=cuda, seed=0)
    rs = np.random.RandomState(0)
    a = torch.from_numpy(rs.randint(0, 256, size=(8, 224, 224, 3), dtype=np.uint8)).to(cuda)
    b = torch.from_numpy(rs.randint(0, 256, size=(8, 224, 224, 3), dtype=np.uint8)).to(cuda)
    run = model.graphed(a)
    la = run(a).clone()
    lb = run(b).clone()
    assert torch.equal(la, model(a)) and torch.equal(lb, model(b))
    labels = la.argmax(1)
    c = torch.zeros(3, dtype=torch.int64, device=cuda)
    ops.topk_count_(c, la, labels)
    assert c.tolist() == [8, 8, 8]


def test_full_batch_256(cuda):
    """BASELINE config 2 size: batch 256 through ResNet-50; batch-invariance of the logits (every
    image is independent) is the size-independent property."""
    from robustart_b200 import nets
    model = nets.build_model("resnet50", device=cuda, seed=0)
    rs = np.random.RandomState(1)
    imgs = torch.from_numpy(rs.randint(0, 256, size=(256, 224, 224, 3), dtype=np.uint8)).to(cuda)
    full = model(imgs)
    part = model(imgs[100:104].contiguous())
    assert torch.isfinite(full).all()
    assert (full[100:104] - part).abs().max().item() < 1e-5
