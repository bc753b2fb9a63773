"""Per-launch report of one ResNet-50 forward (batch 256): shape, CUDA-event time, achieved TFLOP/s and GB/s and
the per-layer roofline time max(bytes/HBM, passes*flops/bf16 peak).  Writes gpurun_out/layer_report.txt.

  python tools/layer_report.py [batch] [arch] [passes: 16 = fp16 single plane (default), 3 = split-bf16]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustart_b200 import nets, ops  # noqa: E402

PK = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}


def main():
    dev = torch.device("cuda", 0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    arch = sys.argv[2] if len(sys.argv) > 2 else "resnet50"
    passes = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    BPE = 2.0 if passes == 16 else 4.0            # activation bytes per element: one fp16 plane or two bf16 planes
    MMAS = 3 if passes == 3 else 1                # tensor-pipe MMAs per algorithmic product
    model = nets.build_model(arch, device=dev, passes=passes)
    img = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev)
    recs = []
    o_conv, o_lin, o_stem, o_mp, o_ap = ops.conv2d_nhwc, ops.linear, ops.stem_conv7x7_u8, ops.maxpool3x3s2, ops.global_avgpool

    def timed(name, fn, flops, byts):
        def w(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            y = fn(*a, **k)
            e.record()
            f, b = flops(*a, **k), byts(*a, **k)
            recs.append([name(*a, **k), s, e, f, b])
            return y
        return w

    def cshape(x, wgt, *a, **k):
        _, nn, h, w, cin = x.shape
        _, cout, kh, kw, _ = wgt.shape
        st, pd = k.get("stride", 1), k.get("pad", 0)
        return nn, h, w, cin, cout, kh, st, (h + 2 * pd - kh) // st + 1, (w + 2 * pd - kw) // st + 1
    ops.conv2d_nhwc = timed(lambda *a, **k: "conv %dx%d s%d %4d->%4d @%3d res=%d" % (cshape(*a, **k)[5], cshape(*a, **k)[5], cshape(*a, **k)[6], cshape(*a, **k)[3], cshape(*a, **k)[4], cshape(*a, **k)[7], int((a[4] if len(a) > 4 else k.get("res")) is not None)), o_conv,
                            lambda *a, **k: 2.0 * cshape(*a, **k)[0] * cshape(*a, **k)[7] * cshape(*a, **k)[8] * cshape(*a, **k)[4] * cshape(*a, **k)[5] ** 2 * cshape(*a, **k)[3],
                            lambda *a, **k: BPE * (a[0][0].numel() + cshape(*a, **k)[0] * cshape(*a, **k)[7] * cshape(*a, **k)[8] * cshape(*a, **k)[4] * (2 if ((a[4] if len(a) > 4 else k.get("res")) is not None) else 1)))
    ops.linear = timed(lambda x, w, *a, **k: "linear %d x %d -> %d" % (x[0].numel() // x.shape[-1], x.shape[-1], w.shape[1]), o_lin,
                       lambda x, w, *a, **k: 2.0 * x[0].numel() * w.shape[1], lambda x, w, *a, **k: BPE * (x[0].numel() + x[0].numel() // x.shape[-1] * w.shape[1]))
    ops.stem_conv7x7_u8 = timed(lambda img, *a, **k: "stem 7x7 s2 u8 fused", o_stem, lambda img, *a, **k: 2.0 * img.shape[0] * 112 * 112 * 64 * 147,
                                lambda img, *a, **k: img.numel() + BPE * img.shape[0] * 112 * 112 * 64)
    ops.stem_pool_u8 = timed(lambda img, *a, **k: "stem 7x7 s2 + maxpool, one launch", ops.stem_pool_u8, lambda img, *a, **k: 2.0 * img.shape[0] * 112 * 112 * 64 * 147,
                             lambda img, *a, **k: img.numel() + BPE * img.shape[0] * 56 * 56 * 64)
    ops.stem_pool_u8_split = timed(lambda img, *a, **k: "stem 7x7 s2 + maxpool, one launch (split)", ops.stem_pool_u8_split, lambda img, *a, **k: 2.0 * img.shape[0] * 112 * 112 * 64 * 147,
                                   lambda img, *a, **k: img.numel() + BPE * img.shape[0] * 56 * 56 * 64)
    ops.maxpool3x3s2 = timed(lambda x, *a, **k: "maxpool", o_mp, lambda x, *a, **k: 0.0, lambda x, *a, **k: BPE * x[0].numel() * 1.25)
    ops.global_avgpool = timed(lambda x, *a, **k: "avgpool", o_ap, lambda x, *a, **k: 0.0, lambda x, *a, **k: BPE * x[0].numel())
    for _ in range(3):
        recs.clear()
        model.forward(img)
        torch.cuda.synchronize()
    lines, tot, tot_roof = [], 0.0, 0.0
    for name, s, e, f, b in recs:
        ms = s.elapsed_time(e)
        roof = max(b / (PK["hbm_gbs"] * 1e9), MMAS * f / (PK["bf16_tflops_sustained"] * 1e12)) * 1e3
        tot += ms
        tot_roof += roof
        lines.append("%-44s %8.3f ms  %7.1f TF/s  %7.1f GB/s  roofline %6.3f ms  (x%.1f)" % (name, ms, f / ms / 1e9, b / ms / 1e6, roof, ms / max(roof, 1e-9)))
    lines.append("TOTAL %.3f ms   roofline-sum %.3f ms   images/s at roofline %.0f" % (tot, tot_roof, n / tot_roof * 1e3))
    os.makedirs("gpurun_out", exist_ok=True)
    lines.insert(0, "%s batch %d passes %d (%s)" % (arch, n, passes, "fp16 single plane" if passes == 16 else "split-bf16"))
    open("gpurun_out/layer_report_p%d.txt" % passes, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
