"""Golden logits from the REFERENCE's own model classes (CPU fp32), this container only.

    python tests/golden/make_golden_models.py

Imports prototype.prototype.model.resnet_official from /root/reference behind import shims (easydict and
a few absent third-party modules that the package __init__ files pull in), loads the deterministic
synthetic weights robustart_b200.nets.random_state_dict(spec, seed=0) with strict=True (this also pins
the key/shape spec against the reference's state_dict), runs 4 synthetic images through
Resize-free ToTensor+Normalize -> model.eval() and stores the logits.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def shim():
    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __setattr__ = __setitem__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)
    m = types.ModuleType("easydict")
    m.EasyDict = EasyDict
    sys.modules["easydict"] = m
    for name in ("foolbox", "art", "art.estimators", "art.estimators.classification", "art.attacks", "art.attacks.evasion",
                 "ffmpeg", "prettytable", "tensorboardX", "timm", "timm.models", "timm.models.layers"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["prettytable"].PrettyTable = object
    sys.modules["tensorboardX"].SummaryWriter = object
    if not hasattr(np, "int"):
        np.int = int


def reference_model(arch):
    # make `prototype` importable from the reference WITHOUT shadowing by this repo's own `prototype`
    # package: load the module file directly.
    import importlib.util
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "prototype" or k.startswith("prototype.")]:
        del sys.modules[k]
    import prototype.prototype.model.resnet_official as R   # noqa
    assert R.__file__.startswith(REF), R.__file__
    return getattr(R, arch + "_official")()


def main():
    from util import synth_images
    from robustart_b200 import nets
    shim()
    images = synth_images(4, seed=7)
    x = torch.from_numpy(images).permute(0, 3, 1, 2).float().div(255)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    xn = (x - mean) / std
    out = {}
    for arch in ("resnet18", "resnet50"):
        model = reference_model(arch)
        spec = nets.resnet_spec(arch)
        ref_sd = model.state_dict()
        assert [(k, tuple(v.shape)) for k, v in ref_sd.items()] == [(k, tuple(s)) for k, s in spec], "spec mismatch"
        sd = nets.random_state_dict(spec, seed=0)
        model.load_state_dict(sd, strict=True)
        model.eval()
        with torch.no_grad():
            logits = model(xn)
        out[arch] = logits.numpy()
        print(arch, "logits abs max %.3f std %.3f" % (logits.abs().max(), logits.std()), "top1", logits.argmax(1).tolist())
    np.savez(os.path.join(HERE, "resnet_logits.npz"), **out)


if __name__ == "__main__":
    main()


def token_models():
    """ViT-B/16 (representation_size 768, model_config.py:216-225) and MLP-Mixer-B/16 golden logits."""
    from util import synth_images
    from robustart_b200 import nets
    shim()
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "prototype" or k.startswith("prototype.")]:
        del sys.modules[k]
    import prototype.prototype.model.vision_transformer as V
    import prototype.prototype.model.vit.mlp_mixer as MX
    assert V.__file__.startswith(REF) and MX.__file__.startswith(REF)
    images = synth_images(2, seed=9)
    x = torch.from_numpy(images).permute(0, 3, 1, 2).float().div(255)
    xn = (x - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    out = {}
    for name, model, spec in (("vit_b16_224", V.vit_b16_224(drop_path=0.0, dropout=0.0, attention_dropout=0.0, qkv_bias=True,
                                                            representation_size=768), nets.vit_spec()),
                              ("mixer_b16_224", MX.mixer_b16_224(), nets.mixer_spec())):
        ref_sd = model.state_dict()
        assert [(k, tuple(v.shape)) for k, v in ref_sd.items()] == [(k, tuple(s)) for k, s in spec], name
        model.load_state_dict(nets.random_token_state_dict(spec, 0), strict=True)
        model.eval()
        with torch.no_grad():
            lg = model(xn)
        out[name] = lg.numpy()
        print(name, "logits abs max %.3f std %.3f" % (lg.abs().max(), lg.std()), lg.argmax(1).tolist())
    np.savez(os.path.join(HERE, "token_logits.npz"), **out)


if __name__ == "__main__" and "--tokens" in sys.argv:
    token_models()


def mobile_models():
    """MobileNetV2 x1.0 and EfficientNet-B0 golden logits from the reference classes."""
    from util import synth_images
    from robustart_b200 import nets
    shim()
    sys.path.insert(0, REF)
    for k in [k for k in sys.modules if k == "prototype" or k.startswith("prototype.")]:
        del sys.modules[k]
    import importlib
    importlib.import_module("prototype.prototype.model.mobilenet_v2")
    importlib.import_module("prototype.prototype.model.efficientnet")
    MB = sys.modules["prototype.prototype.model.mobilenet_v2"]
    EF = sys.modules["prototype.prototype.model.efficientnet"]
    assert MB.__file__.startswith(REF) and EF.__file__.startswith(REF)
    images = synth_images(2, seed=11)
    x = torch.from_numpy(images).permute(0, 3, 1, 2).float().div(255)
    xn = (x - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    out = {}
    for name, model, spec in (("mobilenet_v2", MB.mobilenet_v2(), nets.mobilenet_v2_spec()),
                              ("efficientnet_b0", EF.efficientnet_b0(), nets.efficientnet_b0_spec())):
        ref_sd = model.state_dict()
        a = [(k, tuple(v.shape)) for k, v in ref_sd.items()]
        b = [(k, tuple(s)) for k, s in spec]
        assert a == b, (name, [x for x in zip(a, b) if x[0] != x[1]][:5], len(a), len(b))
        model.load_state_dict(nets.random_state_dict(spec, 0), strict=True)
        model.eval()
        with torch.no_grad():
            lg = model(xn)
        out[name] = lg.numpy()
        print(name, "logits abs max %.3f std %.3f" % (lg.abs().max(), lg.std()), lg.argmax(1).tolist())
    np.savez(os.path.join(HERE, "mobile_logits.npz"), **out)


if __name__ == "__main__" and "--mobile" in sys.argv:
    mobile_models()
