"""Per-op CUDA-event time of one eager forward (batch given) for any model family: python tools/op_times.py arch [batch] [passes]
Wraps every public function of robustart_b200.ops that a forward calls; prints the share of each op name (with the
activation / shape class for linear and conv)."""
import collections, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import nets, ops

arch = sys.argv[1]
args = [a for a in sys.argv[1:] if not a.startswith("--")]
arch = args[0]
n = int(args[1]) if len(args) > 1 else 128
passes = int(args[2]) if len(args) > 2 else 3
dev = torch.device("cuda", 0)
build = getattr(nets, "build_any", None) or nets.build_model
try:
    model = build(arch, device=dev, passes=passes)
except Exception:
    from robustart_b200 import solver
    model = solver.build_native_model(arch, dev, passes=passes) if hasattr(solver, "build_native_model") else None
    if model is None:
        raise
img = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev)
recs = []
names = [k for k, v in vars(ops).items() if callable(v) and not k.startswith("_") and getattr(v, "__module__", "") == ops.__name__]
orig = {k: getattr(ops, k) for k in names}
depth = [0]


def wrap(name, fn):
    def w(*a, **k):
        if depth[0]:
            return fn(*a, **k)
        depth[0] += 1
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        try:
            y = fn(*a, **k)
        finally:
            depth[0] -= 1
        e.record()
        tag = name
        if name in ("linear", "conv2d_nhwc"):
            x, wg = a[0], a[1]
            tag += " act=%s res=%d" % (k.get("act"), int((a[4] if len(a) > 4 else k.get("res")) is not None))
            if name == "linear":
                tag += " k=%d n=%d" % (x.shape[-1], wg.shape[1])
            else:
                tag += " %dx%d c%d->%d s%d @%d" % (wg.shape[2], wg.shape[3], x.shape[-1], wg.shape[1], k.get("stride", 1), x.shape[2])
        if name == "conv2d_dgrad":
            dy, wt = a[0], a[1]
            tag += " %dx%d c%d->%d @%d res=%d mask=%d" % (wt.shape[2], wt.shape[3], dy.shape[-1], wt.shape[1], dy.shape[2],
                                                        int((a[2] if len(a) > 2 else k.get("res")) is not None), int((a[3] if len(a) > 3 else k.get("mask")) is not None))
        recs.append((tag, s, e))
        return y
    return w


for k in names:
    setattr(ops, k, wrap(k, orig[k]))
grad = "--grad" in sys.argv          # forward_saved + input gradient (one PGD step's model work) instead of the forward
if grad:
    x01 = torch.rand(n, 3, 224, 224, device=dev)
    yl = torch.randint(0, 1000, (n,), device=dev)
for _ in range(3):
    recs.clear()
    if grad:
        lg, saved = model.forward_saved(x01)
        model.input_grad(ops.ce_loss_grad(lg, yl)[1], saved)
    else:
        model.forward(img)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for tag, s, e in recs:
    a = agg.setdefault(tag, [0, 0.0])
    a[0] += 1
    a[1] += s.elapsed_time(e)
tot = sum(v[1] for v in agg.values())
print("%s batch %d passes %d: %.3f ms per forward, %.0f img/s (eager, event per op)" % (arch, n, passes, tot, n / tot * 1e3))
for tag, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %6.3f ms %5.1f%%  x%-3d %s" % (ms, 100 * ms / tot, c, tag))
