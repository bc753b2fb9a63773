import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from robustart_b200 import ops, nets, torch_models
from util import diverse_images, calibrated_state_dict
dev = torch.device("cuda", 0)
torch.manual_seed(0)
# 1. accumulation bias probe: positive operands, long K
for K in (512, 2048, 8192):
    for passes in (3, 1):
        x = torch.rand(512, K, device=dev) + 0.5
        w = torch.rand(128, K, device=dev) + 0.5
        xp, wp = ops.split_f32(x), ops.split_f32(w)
        xr = ops.merge_f32(xp) if passes == 3 else ops.from_planes(xp[:1].contiguous())
        wr = ops.merge_f32(wp) if passes == 3 else ops.from_planes(wp[:1].contiguous())
        out = torch.empty(512, 128, device=dev)
        ops.linear(xp, wp, None, None, passes=passes, out_f32=out, want_planes=False)
        ref = xr.double() @ wr.double().t()
        rel = ((out.double() - ref) / ref)
        print("K=%5d passes=%d  signed mean rel err %+.3e   rms %.3e  max %.3e" % (K, passes, rel.mean().item(), rel.pow(2).mean().sqrt().item(), rel.abs().max().item()))
# 2. per-block error of calibrated ResNet-50 against the fp64 twin
cal = np.load("tests/golden/calibrated_logits.npz")
for arch in ("resnet50", "resnet18"):
    sd = calibrated_state_dict(arch, nets.random_state_dict(nets.resnet_spec(arch), 0), cal)
    net = nets.build_model(arch, sd, device=dev, passes=3)
    twin = torch_models.build(arch, sd).to(dev).double().eval()
    imgs = torch.from_numpy(diverse_images(8, seed=0)).to(dev)
    x01 = imgs.permute(0, 3, 1, 2).float().div(255).contiguous()
    logits, saved = net.forward_saved(x01)
    mean = torch.tensor(ops.IMAGENET_MEAN, device=dev, dtype=torch.float64).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD, device=dev, dtype=torch.float64).view(1, 3, 1, 1)
    feats = []
    hooks = []
    for name, m in twin.named_modules():
        if name.count(".") == 1 and name.startswith("layer"):
            hooks.append(m.register_forward_hook(lambda mod, i, o, name=name: feats.append((name, o.detach()))))
    with torch.no_grad():
        ref_logits = twin((x01.double() - mean) / std)
    print(arch, "logits err vs fp64 twin: %.3e ; golden vs twin %.3e" % ((logits.double() - ref_logits).abs().max().item(), np.abs(cal[arch + "/logits"] - ref_logits.cpu().numpy()).max()))
    for (name, f), sv in zip(feats, saved["blocks"]):
        y = ops.merge_f32(sv[-1]).permute(0, 3, 1, 2).double()
        d = (y - f)
        print("  %-10s rel rms err %.3e   signed mean of err/|f|mean %+.3e" % (name, (d.pow(2).mean().sqrt() / f.pow(2).mean().sqrt()).item(), (d.mean() / f.abs().mean()).item()))
    # u8 path
    lg8 = net.forward(imgs)
    print("  u8-path logits err vs twin %.3e, vs golden %.3e" % ((lg8.double() - ref_logits).abs().max().item(), np.abs(lg8.cpu().numpy() - cal[arch + "/logits"]).max()))
