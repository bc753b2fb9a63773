import torch, sys, json
sys.path.insert(0, '.')
from robustart_b200 import nets, ops
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
ins = [torch.randint(0, 256, (256, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(4)]
for P in (3, 16):
    for arch in ("resnet50", "resnet18"):
        m = nets.build_model(arch, device=dev, seed=0, passes=P)
        run = m.graphed(ins[0])
        for i in range(3): run(ins[i % 4])
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(10): run(ins[i % 4])
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print(json.dumps({"arch": arch, "passes": P, "ms_per_256": ms, "img_per_s": 256 / ms * 1e3}), flush=True)
