"""Realistic-magnitude golden logits from the REFERENCE's own model classes (CPU fp32), this container only.

    python tests/golden/make_golden_calibrated.py

VERDICT r1 (weak #1): the synthetic weights give logits with |max| ~0.5, so an absolute 1e-3 bar on them says little about
a trained network (|logit| 10-30).  Here the classifier head of every BASELINE family is calibrated on 8 structurally
different images (tests/util.diverse_images): head.weight *= s, head.bias = -0.9 s W mean(features), with s chosen so that
the logits have std 2.5 -- a trained ImageNet classifier's scale -- and the images get different top-1 classes.  The
calibration (s, bias) and the logits the reference's class then produces are stored; tests rebuild the same state_dict
with tests/util.calibrated_state_dict.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_models import shim, REF, ROOT  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))


def ref_models():
    shim()
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "prototype" or k.startswith("prototype.")]:
        del sys.modules[k]
    import importlib
    R = importlib.import_module("prototype.prototype.model.resnet_official")
    V = importlib.import_module("prototype.prototype.model.vision_transformer")
    MX = importlib.import_module("prototype.prototype.model.vit.mlp_mixer")
    MB = importlib.import_module("prototype.prototype.model.mobilenet_v2")
    EF = importlib.import_module("prototype.prototype.model.efficientnet")
    for m in (R, V, MX, MB, EF):
        assert m.__file__.startswith(REF), m.__file__
    from robustart_b200 import nets
    return [
        ("resnet18", R.resnet18_official(), nets.random_state_dict(nets.resnet_spec("resnet18"), 0)),
        ("resnet50", R.resnet50_official(), nets.random_state_dict(nets.resnet_spec("resnet50"), 0)),
        ("vit_b16_224", V.vit_b16_224(drop_path=0.0, dropout=0.0, attention_dropout=0.0, qkv_bias=True, representation_size=768),
         nets.random_token_state_dict(nets.vit_spec(), 0)),
        ("mixer_b16_224", MX.mixer_b16_224(), nets.random_token_state_dict(nets.mixer_spec(), 0)),
        ("mobilenet_v2", MB.mobilenet_v2(), nets.random_state_dict(nets.mobilenet_v2_spec(), 0)),
        ("efficientnet_b0", EF.efficientnet_b0(), nets.random_state_dict(nets.efficientnet_b0_spec(), 0)),
    ]


# share of the mean feature the head's bias cancels.  0.9 for the families whose features vary by >= 50 % between images; the
# MobileNetV2 random-weight features vary by 20 % only, and cancelling 90 % of the mean there would make the logits a
# small difference of large numbers (|s W f| of several hundred at logit std 2.5), which no trained network is
CENTRE = {"mobilenet_v2": 0.6}


def main():
    from util import diverse_images, calibrated_state_dict, HEAD_KEYS
    images = diverse_images(8, seed=0)
    x = torch.from_numpy(images).permute(0, 3, 1, 2).float().div(255)
    xn = (x - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    out = {}
    for arch, model, sd in ref_models():
        model.load_state_dict(sd, strict=True)
        cal = {}
        bns = [(n, m) for n, m in model.named_modules() if isinstance(m, torch.nn.BatchNorm2d)]
        if bns:
            # what training leaves in the BN buffers: the statistics of the data.  momentum 1 -> running = batch statistics
            model.train()
            for _, m in bns:
                m.momentum = 1.0
            with torch.no_grad():
                model(xn)
            sd = dict(sd)
            for n, m in bns:
                for b in ("running_mean", "running_var"):
                    cal["%s/bn/%s.%s" % (arch, n, b)] = getattr(m, b).detach().float().numpy().copy()
                    sd["%s.%s" % (n, b)] = getattr(m, b).detach().clone()
            model.load_state_dict(sd, strict=True)
        model.eval()
        head = model
        for part in HEAD_KEYS[arch].split("."):
            head = getattr(head, part) if not part.isdigit() else head[int(part)]
        feats = []
        h = head.register_forward_hook(lambda m, i, o: feats.append(i[0].detach().clone()))
        with torch.no_grad():
            model(xn)
        h.remove()
        f = feats[0].double()
        W = head.weight.detach().double()
        c = CENTRE.get(arch, 0.9)
        centred = (f - c * f.mean(0, keepdim=True)) @ W.t()
        s = np.float32(2.5 / centred.std().item())
        bias = (-(c * float(s)) * (W @ f.mean(0))).float().numpy()
        cal.update({arch + "/scale": s, arch + "/bias": bias})
        model.load_state_dict(calibrated_state_dict(arch, sd, cal), strict=True)
        with torch.no_grad():
            lg = model(xn)
        out.update(cal)
        out[arch + "/logits"] = lg.numpy()
        print("%-16s scale %.3f  logits |max| %.2f std %.2f  top1 %s  (feature variation %.2f of |f|, |s W f| max %.1f)" % (
            arch, s, lg.abs().max(), lg.std(), lg.argmax(1).tolist(), ((f - f.mean(0)).norm() / f.norm()).item(),
            (float(s) * (f @ W.t())).abs().max().item()))
    np.savez_compressed(os.path.join(HERE, "calibrated_logits.npz"), **out)


if __name__ == "__main__":
    main()
