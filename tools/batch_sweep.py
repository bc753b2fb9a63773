"""images/s of the graph-replayed forward vs batch size (L2 residency of the activations vs tile-count effects)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import nets
dev = torch.device("cuda", 0)
arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
m = nets.build_model(arch, device=dev, seed=0, passes=passes)
for n in (8, 16, 32, 64, 128, 256):
    ins = [torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(max(2, 512 // n))]
    run = m.graphed(ins[0])
    for i in range(3): run(ins[i % len(ins)])
    torch.cuda.synchronize()
    reps = max(8, 1024 // n)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps): run(ins[i % len(ins)])
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(json.dumps({"arch": arch, "batch": n, "ms": round(ms, 4), "img_per_s": round(n / ms * 1e3)}), flush=True)
