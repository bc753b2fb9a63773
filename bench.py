#!/usr/bin/env python
"""Headline benchmark: BASELINE.json configs[1] -- ResNet-50 + ImageNet-C gaussian_noise severity 1-5,
batch 256 per GPU.  One "step" = one pass of the hot path over one batch of 256 synthetic 224x224x3 uint8
images: AddNoise gaussian_noise (severity cycles 1..5) -> normalise -> ResNet-50 forward -> top-1/top-5
counters.  metric = corrupted-images/sec (whole job, all GPUs).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels)
  python bench.py --impl reference ...                           the reference's CPU path on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...              N > 1: one rank per GPU, weak scaling

Prints ONE JSON line (contract in the task statement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 256
H = W = 224
R_INPUTS = 8      # rotating input batches: 8 x 38.5 MB = 308 MB > 126 MB L2
METRIC = "corrupted-images/sec"
WORKLOAD = "configs[1]: ResNet-50 + ImageNet-C gaussian_noise severity 1-5 (cycled per step), batch 256/GPU"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, str(gpu_index)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.gpu], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU path restated in oracle/ (numpy AddNoise loop +
# torch fp32 ResNet-50 on the host cores).  This is the only place bench.py executes oracle/.
# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(n_images, threads, seed=0):
    import numpy as np
    import torch
    from oracle import imagenet_c as O
    from oracle import models as OM
    from oracle import metrics as OMet
    from robustart_b200 import nets
    torch.set_num_threads(threads)
    if not hasattr(cpu_reference_pass, "_model"):
        cpu_reference_pass._model = OM.build("resnet50", nets.random_state_dict(nets.resnet_spec("resnet50"), 0))
    model = cpu_reference_pass._model
    rs = np.random.RandomState(seed)
    images = rs.randint(0, 256, size=(n_images, H, W, 3), dtype=np.uint8)
    labels = torch.from_numpy(rs.randint(0, 1000, size=(n_images,)))
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    t0 = time.perf_counter()
    O.add_noise_for_imagenet_c(images, severity=1 + seed % 5, corruption_name="gaussian_noise", draws=O.NumpyDraws(seed))
    t1 = time.perf_counter()
    with torch.no_grad():
        x = (torch.from_numpy(images).permute(0, 3, 1, 2).float().div(255) - mean) / std
        logits = model(x)
    hits = OMet.topk_hits(logits, labels)
    t2 = time.perf_counter()
    return {"total_s": t2 - t0, "noise_s": t1 - t0, "model_s": t2 - t1, "hits": hits}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = 32
    for _ in range(max(args.warmup, 1)):
        cpu_reference_pass(per_step, threads)
    t = 0.0
    for i in range(args.steps):
        t += cpu_reference_pass(per_step, threads, seed=i)["total_s"]
    value = per_step * args.steps / t
    sample = "%d steps x %d images of the same workload (numpy gaussian_noise loop + torch fp32 ResNet-50)" % (args.steps, per_step)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reference_batch_per_step": per_step,
                       "note": "reference CPU path = oracle port (the reference package itself cannot be imported: skimage/wand/easydict absent)"},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
DTYPE = {"f16": "f16 (fp16 operands and activations, fp32 accumulation in TMEM; logits within 1e-3 of the fp32 reference)",
         "bf16x3": "bf16x3 (split-bf16 operands, fp32 accumulate: fp32-faithful)"}


def small_kernels_ms(model, pipe, reps=20):
    """CUDA-event time of the non-tensor-core kernels of a step besides the corruption: global average pool + counters."""
    import torch
    from robustart_b200 import ops
    n = pipe.static_in.shape[0]
    f16 = getattr(model, "f16", False)
    c = model.fc_w.shape[-1]
    feat = torch.zeros((1 if f16 else 2, n, 7, 7, c), dtype=torch.int16, device=pipe.device)
    logits = torch.randn(n, model.num_classes, device=pipe.device)
    labels = torch.zeros(n, dtype=torch.int64, device=pipe.device)
    counters = torch.zeros(3, dtype=torch.int64, device=pipe.device)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.global_avgpool(feat); ops.topk_count_(counters, logits, labels, pipe.pred)
    s.record()
    for _ in range(reps):
        ops.global_avgpool(feat); ops.topk_count_(counters, logits, labels, pipe.pred)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def gemm_roofline(model, pipe, pk, precision):
    """Live per-launch timing of the dominant kernel (tcgen05 implicit GEMM) with CUDA events on the
    launching stream: the forward is re-run eagerly with an event pair around every conv / linear."""
    import torch
    from robustart_b200 import ops
    recs = []
    orig_conv, orig_lin = ops.conv2d_nhwc, ops.linear

    def conv(x, wgt, *a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = orig_conv(x, wgt, *a, **k)
        e.record()
        _, n, h, w, cin = x.shape
        _, cout, kh, kw, _ = wgt.shape
        st, pd = k.get("stride", 1), k.get("pad", 0)
        ho, wo = (h + 2 * pd - kh) // st + 1, (w + 2 * pd - kw) // st + 1
        has_res = (a[2] if len(a) > 2 else k.get("res")) is not None
        bpe = 2.0 if x.shape[0] == 1 else 4.0            # one fp16 plane or two bf16 planes per element
        byts = bpe * (n * h * w * cin + n * ho * wo * cout * (2 if has_res else 1))
        recs.append((s, e, 2.0 * n * ho * wo * cout * kh * kw * cin, byts))
        return y

    def lin(x, wgt, *a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = orig_lin(x, wgt, *a, **k)
        e.record()
        kk = x.shape[-1]
        rows = x[0].numel() // kk
        recs.append((s, e, 2.0 * rows * kk * wgt.shape[1], (2.0 if x.shape[0] == 1 else 4.0) * (rows * kk + rows * wgt.shape[1])))
        return y

    orig_stem = ops.stem_conv7x7_u8

    def stem(img, *a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = orig_stem(img, *a, **k)
        e.record()
        recs.append((s, e, 2.0 * img.shape[0] * (img.shape[1] // 2) * (img.shape[2] // 2) * 64 * 147,
                     img.numel() + (2.0 if a[0].shape[0] == 1 else 4.0) * img.shape[0] * (img.shape[1] // 2) * (img.shape[2] // 2) * 64))
        return y

    orig_stem_pool = ops.stem_pool_u8

    def stem_pool(img, *a, **k):       # one-launch stem (conv1 + bn + relu + maxpool on the tensor core, fp16 mode)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = orig_stem_pool(img, *a, **k)
        e.record()
        recs.append((s, e, 2.0 * img.shape[0] * (img.shape[1] // 2) * (img.shape[2] // 2) * 64 * 147,
                     img.numel() + 2.0 * img.shape[0] * (img.shape[1] // 4) * (img.shape[2] // 4) * 64))
        return y

    ops.conv2d_nhwc, ops.linear, ops.stem_conv7x7_u8, ops.stem_pool_u8 = conv, lin, stem, stem_pool
    try:
        for _ in range(2):
            recs.clear()
            model.forward(pipe.static_in)
            torch.cuda.synchronize()
    finally:
        ops.conv2d_nhwc, ops.linear, ops.stem_conv7x7_u8, ops.stem_pool_u8 = orig_conv, orig_lin, orig_stem, orig_stem_pool
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_gemm_traffic_%s.json" % precision)
    if os.path.exists(tp):      # dram__bytes_read+write summed over the same launches, from one ncu capture (profiles/)
        traffic = json.load(open(tp)).get("dram_bytes_per_forward")
    ms = sum(s.elapsed_time(e) for s, e, _, _ in recs)
    flops = sum(f for _, _, f, _ in recs)
    achieved = flops / (ms * 1e-3) / 1e12
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    # per-launch roofline: a layer can go no faster than its activation traffic over HBM or its MMAs over the tensor pipe
    mmas = 1.0 if precision == "f16" else 3.0
    layer_ms = sum(max(b / (pk["hbm_gbs"] * 1e9), mmas * f / (peak * 1e12)) for _, _, f, b in recs) * 1e3
    return {"bound": "tensor", "kernel": "gemm_kernel<BN> + stem_pool_kernel (tcgen05 implicit GEMMs, all %d launches of one forward)" % len(recs),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_note": "bytes per forward (all GEMM launches), ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_gemm_traffic_%s.json" % precision,
            "peak_source": pk["_source"] + ", sustained bf16 (kernel timed inside a long step)",
            "gemm_ms_per_step": ms, "algorithmic_gflop_per_step": flops / 1e9,
            "layer_roofline_ms_per_step": layer_ms, "frac_of_layer_roofline": layer_ms / ms,
            "layer_roofline_note": "sum over the launches of max(algorithmic activation bytes / HBM peak, issued MMA FLOPs / tensor peak): "
                                   "most 1x1 layers of ResNet-50 are HBM-bound, so the tensor-only `frac` cannot approach 1",
            "note": ("algorithmic FLOPs = 2*M*N*K of the convolution, one fp16 MMA per product; most 1x1 layers of ResNet-50 are bound by "
                     "activation traffic, not the tensor pipe (profiles/ layer report)") if precision == "f16" else
                    ("algorithmic FLOPs = 2*M*N*K of the fp32 convolution; the split-bf16 scheme issues 3 bf16 MMAs per "
                     "product, so tensor-pipe FLOP/s are 3x achieved")}


def corruption_roofline(pipe, inputs, pk):
    """Kernel-only time of the gaussian_noise launch at the bench batch (256 images = 77 MB of traffic, ~12 us at
    HBM speed -- less than a Python launch), so the launches are captured into a CUDA graph (one per rotating
    input batch) and the graph is replayed: CUDA-event time / launches = device time per launch."""
    import torch
    from robustart_b200 import ops
    outs = [torch.empty_like(inputs[0]) for _ in range(len(inputs))]
    for i in range(3):
        ops.corrupt_u8(inputs[i % len(inputs)], "gaussian_noise", 1 + i % 5, seed=i, out=outs[i % len(outs)])
    torch.cuda.synchronize()
    per_graph = 2 * len(inputs)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(per_graph):
                ops.corrupt_u8(inputs[i % len(inputs)], "gaussian_noise", 1 + i % 5, seed=i, out=outs[i % len(outs)])
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    t = s.elapsed_time(e) * 1e-3 / (reps * per_graph)
    alg = 2.0 * BATCH * H * W * 3
    return {"bound": "hbm", "kernel": "normal_noise_kernel (gaussian_noise, u8 NHWC -> u8 NHWC)", "achieved": alg / t / 1e9,
            "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": alg / t / 1e9 / pk["hbm_gbs"], "traffic": None,
            "us_per_launch": t * 1e6, "algorithmic_bytes_per_launch": alg, "images_per_s_kernel_only": BATCH / t,
            "timing": "CUDA graph of %d launches over %d rotating input batches (> L2), %d replays" % (per_graph, len(inputs), reps)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from robustart_b200 import nets, _lib
    from robustart_b200.evalpipe import CorruptEvalPipeline

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()

    passes = {"f16": 16, "bf16x3": 3}[args.precision]
    model = nets.build_model("resnet50", device=dev, seed=0, passes=passes)
    pipe = CorruptEvalPipeline(model, BATCH, H, W, seed=1234 + rank)
    g = torch.Generator(device=dev).manual_seed(rank)
    inputs = [torch.randint(0, 256, (BATCH, H, W, 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(R_INPUTS)]
    labels = torch.randint(0, 1000, (BATCH,), device=dev, generator=g)
    h_inputs = [t.cpu().pin_memory() for t in inputs[:4]]
    h_labels = labels.cpu().pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ---------------------------------------------------
    for i in range(args.warmup):
        pipe.step_device(inputs[i % R_INPUTS], labels, "gaussian_noise", 1 + i % 5)
    pipe.reset()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        pipe.step_device(inputs[i % R_INPUTS], labels, "gaussian_noise", 1 + i % 5)
    e.record()
    barrier()
    dev_ms = s.elapsed_time(e)
    clocks = sampler.stop() if rank == 0 else None
    pipe.allreduce_counters()                      # the one collective of the evaluation
    counts = pipe.counters.tolist()

    # ---- end-to-end through host buffers ("e2e") ----------------------------------------------
    for i in range(max(3, args.warmup // 2)):
        pipe.step_host_pipelined(h_inputs[i % len(h_inputs)], h_labels, "gaussian_noise", 1 + i % 5)
    pipe.finish()
    barrier()
    t0 = time.perf_counter()
    s.record()
    for i in range(args.steps):
        pipe.step_host_pipelined(h_inputs[i % len(h_inputs)], h_labels, "gaussian_noise", 1 + i % 5)
    host_counters = pipe.finish()          # the last step's counters are on the host before the clock stops
    e.record()
    barrier()
    e2e_ms = max(s.elapsed_time(e), 0.0)
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()

    if rank == 0:
        value = world * BATCH * args.steps / (dev_ms * 1e-3)
        e2e_v = world * BATCH * args.steps / (e2e_ms * 1e-3)
        roof = gemm_roofline(model, pipe, pk, args.precision)
        roof_c = corruption_roofline(pipe, inputs, pk)
        # the same launches inside the timed region (CUDA-graph replay, no launch gaps): step time minus the step's other
        # kernels (corruption, avgpool, counters -- each timed alone with events)
        other_ms = roof_c["us_per_launch"] * 1e-3 + small_kernels_ms(model, pipe)
        in_graph_ms = dev_ms / args.steps - other_ms
        if in_graph_ms > 0:
            roof["gemm_ms_per_step_in_timed_region"] = in_graph_ms
            roof["achieved_in_timed_region"] = roof["algorithmic_gflop_per_step"] / in_graph_ms      # GFLOP / ms = TFLOP/s
            roof["frac_in_timed_region"] = roof["achieved_in_timed_region"] / roof["peak"]
            roof["frac_of_layer_roofline_in_timed_region"] = roof["layer_roofline_ms_per_step"] / in_graph_ms
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu_reference_pass(16, threads)
            r = [cpu_reference_pass(32, threads, seed=i) for i in range(3)]
            tot = sum(x["total_s"] for x in r)
            cpu = {"value": 96 / tot, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": "3 x 32 images of the same workload: oracle numpy gaussian_noise loop (%.0f img/s alone) + torch fp32 "
                             "ResNet-50 forward (%.0f img/s alone)" % (96 / sum(x["noise_s"] for x in r), 96 / sum(x["model_s"] for x in r))}
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": DTYPE[args.precision], "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": world * BATCH, "image": "224x224x3 uint8 NHWC",
                           "weights": "synthetic (nets.random_state_dict seed 0)", "parallelism": "dp%d" % world,
                           "l2": "inputs rotate over %d batches = %d MB > 126 MB L2" % (R_INPUTS, R_INPUTS * BATCH * H * W * 3 // 2 ** 20),
                           "forward": "CUDA graph replay", "counters": [int(c) for c in counts]},
                "clocks": clocks,
                "e2e": {"value": e2e_v, "unit": "images/s", "h2d_bytes_per_step": BATCH * H * W * 3 + BATCH * 8,
                        "d2h_bytes_per_step": 24, "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall_ms / args.steps,
                        "api": "CorruptEvalPipeline.step_host_pipelined + finish (pinned host uint8 batch -> H2D on a copy stream -> device step -> "
                               "counters D2H every step; the host waits for step i-1's counters while step i runs)",
                        "counters_on_host": [int(v) for v in host_counters.tolist()]},
                "gpu_launches": pipe.launches_per_step * args.steps,
                "gpu_launches_per_step": pipe.launches_per_step,
                "roofline": roof, "roofline_corruption": roof_c}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_pgd:
            line["pgd_loop"] = pgd_loop_report(model, dev, pk, args.precision)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def pgd_loop_report(model, dev, pk, precision, n=128, steps=10, reps=2):
    """The second half of the north star, reported beside the headline metric (not part of `value`): the PGD-Linf
    10-step eval loop on ResNet-50, forward + input gradient on the sm_100a kernels, then one
    forward of the adversarial batch + counters.  Algorithmic FLOPs = (2k+1) * 8.18 GFLOP / image."""
    import torch
    from robustart_b200 import attacks, ops
    g = torch.Generator(device=dev).manual_seed(2)
    x = torch.rand(n, 3, 224, 224, device=dev, generator=g)
    y = torch.randint(0, 1000, (n,), device=dev, generator=g)
    src = attacks.NativeModel(model)
    counters = torch.zeros(3, dtype=torch.int64, device=dev)

    def loop():
        adv = attacks.pgd_linf(x, y, src, 4 / 255, 3 / 40, steps, seed=0)
        ops.topk_count_(counters, model.forward(adv), y)

    loop()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        loop()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    tf = (2 * steps + 1) * 8.18 * n / ms
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    return {"workload": "ResNet-50, pgd_linf eps 4/255, %d steps, batch %d, float32 NCHW images resident in HBM" % (steps, n),
            "images_per_s": n / ms * 1e3, "ms_per_batch": ms, "algorithmic_tflops": tf, "peak_tflops": peak, "frac": tf / peak,
            "source_model": "native dgrad (tcgen05 GEMM on transposed weights, %s), no autograd" % precision,
            "note": "fp16 gradients run loss-scaled (x4096), unscaled when the image gradient is written" if precision == "f16"
                    else "split-bf16 issues 3 MMAs per product: tensor-pipe FLOP/s are 3x algorithmic"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-pgd", action="store_true", help="skip the PGD-loop side report")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16x3"],
                    help="f16: one fp16 plane per tensor, one MMA per product (default); bf16x3: split-bf16, fp32-faithful")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
