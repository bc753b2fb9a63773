"""Autograd-capable PyTorch twins of the kernel models -- used ONLY as the *source* model of an attack
(the input gradient needs a backward pass; dgrad kernels are the next step).  Same
state_dict keys as the reference (resnet_official.py) so one checkpoint feeds both paths."""
import torch
import torch.nn as nn
import torch.nn.functional as F

_CFG = {"resnet18": (False, [2, 2, 2, 2]), "resnet34": (False, [3, 4, 6, 3]), "resnet50": (True, [3, 4, 6, 3]),
        "resnet101": (True, [3, 4, 23, 3])}


class _Block(nn.Module):
    def __init__(self, cin, planes, stride, bottleneck):
        super().__init__()
        cout = planes * (4 if bottleneck else 1)
        if bottleneck:
            shapes = [(cin, planes, 1, 1, 0), (planes, planes, 3, stride, 1), (planes, cout, 1, 1, 0)]
        else:
            shapes = [(cin, planes, 3, stride, 1), (planes, planes, 3, 1, 1)]
        for i, (a, b, k, s, p) in enumerate(shapes, 1):
            setattr(self, "conv%d" % i, nn.Conv2d(a, b, k, s, p, bias=False))
            setattr(self, "bn%d" % i, nn.BatchNorm2d(b))
        self.n = len(shapes)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idn = x if self.downsample is None else self.downsample(x)
        out = x
        for i in range(1, self.n + 1):
            out = getattr(self, "bn%d" % i)(getattr(self, "conv%d" % i)(out))
            if i < self.n:
                out = F.relu(out)
        return F.relu(out + idn)


class ResNet(nn.Module):
    def __init__(self, arch, num_classes=1000):
        super().__init__()
        bott, layers = _CFG[arch]
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, (planes, nb) in enumerate(zip([64, 128, 256, 512], layers), 1):
            blocks = []
            for b in range(nb):
                blocks.append(_Block(cin, planes, 2 if (b == 0 and li > 1) else 1, bott))
                cin = planes * (4 if bott else 1)
            setattr(self, "layer%d" % li, nn.Sequential(*blocks))
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        for li in range(1, 5):
            x = getattr(self, "layer%d" % li)(x)
        return self.fc(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))


def build(arch, state_dict=None):
    m = ResNet(arch)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m.eval()
