#!/usr/bin/env python
"""Headline benchmark: BASELINE.json configs[1] -- ResNet-50 + ImageNet-C gaussian_noise severity 1-5,
batch 256 per GPU.  One "step" = one pass of the hot path over one batch of 256 synthetic 224x224x3 uint8
images: AddNoise gaussian_noise (severity cycles 1..5) -> normalise -> ResNet-50 forward -> top-1/top-5
counters.  metric = corrupted-images/sec (whole job, all GPUs).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels)
  python bench.py --impl reference ...                           the reference's CPU path on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...              N > 1: one rank per GPU, weak scaling

Prints ONE JSON line (contract in the task statement).  Everything printed follows from a measurement made in this
run: `roofline.frac` uses only time spent inside the timed region; the counter all-reduce (the evaluation's one
collective) is inside the timed region; `top1_match` compares the GPU path with the CPU port on the same images
and the same noise draws.
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
RESNET50_GFLOP = 8.18     # SURVEY 8(d), forward, per image (1 MAC = 2 FLOP)


def make_config(world):
    """The SAME dict for both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "global_batch": world * BATCH, "image": "224x224x3 uint8 NHWC",
            "weights": "synthetic, calibrated (tests/golden/calibrated_logits.npz: logit std 2.5, |max| 20, distinct top-1 per image)",
            "parallelism": "dp%d" % world,
            "l2": "every step reads a different input batch: %d rotating batches = %d MB > 126 MB L2" % (
                R_INPUTS, R_INPUTS * BATCH * H * W * 3 // 2 ** 20)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback (B200_PROFILING.md)"}


def resnet50_weights():
    """Synthetic ResNet-50 state_dict: the deterministic random weights with the calibrated head / BN statistics of
    tests/golden/calibrated_logits.npz when that fixture is present (so that predictions differ per image and
    `top1_match` means something); FLOPs and bytes are those of the architecture either way."""
    from robustart_b200 import nets
    sd = nets.random_state_dict(nets.resnet_spec("resnet50"), 0)
    p = os.path.join(ROOT, "tests", "golden", "calibrated_logits.npz")
    if os.path.exists(p):
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from util import calibrated_state_dict
        sd = calibrated_state_dict("resnet50", sd, np.load(p))
    return sd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe), 20 ms period."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, str(gpu_index)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", self.gpu], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 9]
        num = lambda s: s.replace(".", "").isdigit()
        busy = [r for r in rows if num(r[8]) and float(r[8]) > 0] or rows        # samples taken under load
        sm = sorted(int(float(r[1])) for r in busy if num(r[1]))
        mx = [int(float(r[2])) for r in rows if num(r[2])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[4 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in rows if num(r[3])]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(rows), "samples_under_load": len(sm), "power_w_max": max(pw) if pw else None,
                "window": "value loop + e2e loop + kernel-only measurements"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU path restated in oracle/ (numpy AddNoise loop +
# torch fp32 ResNet-50 on the host cores).  This is the only place bench.py executes oracle/.
# ------------------------------------------------------------------------------------------------
def _cpu_model():
    from oracle import models as OM
    if not hasattr(_cpu_model, "m"):
        _cpu_model.m = OM.build("resnet50", resnet50_weights())
    return _cpu_model.m


def cpu_reference_pass(n_images, threads, seed=0, images=None, keep_logits=False):
    """One pass of the reference's CPU path over n_images: the per-image numpy loop of add_noise_utils.py:27-31 with
    gaussian_noise (corruptions.py:122-126), ToTensor + Normalize, torch fp32 ResNet-50, top-k hits."""
    import numpy as np
    import torch
    from oracle import imagenet_c as O
    from oracle import metrics as OMet
    torch.set_num_threads(threads)
    model = _cpu_model()
    rs = np.random.RandomState(seed)
    if images is None:
        images = rs.randint(0, 256, size=(n_images, H, W, 3), dtype=np.uint8)
    labels = torch.from_numpy(rs.randint(0, 1000, size=(n_images,)))
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    draws = O.NumpyDraws(seed)
    t0 = time.perf_counter()
    O.add_noise_for_imagenet_c(images, severity=1 + seed % 5, corruption_name="gaussian_noise", draws=draws)
    t1 = time.perf_counter()
    with torch.no_grad():
        x = (torch.from_numpy(images).permute(0, 3, 1, 2).float().div(255) - mean) / std
        logits = model(x)
    hits = OMet.topk_hits(logits, labels)
    t2 = time.perf_counter()
    out = {"total_s": t2 - t0, "noise_s": t1 - t0, "model_s": t2 - t1, "hits": hits}
    if keep_logits:
        out.update(logits=logits, corrupted=images, draws=draws)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = BATCH
    for _ in range(min(max(args.warmup, 1), 2)):             # warm-up passes are bounded too: a pass is ~3 s of host time
        cpu_reference_pass(64, threads)
    t = 0.0
    for i in range(args.steps):
        t += cpu_reference_pass(per_step, threads, seed=i)["total_s"]
    value = per_step * args.steps / t
    sample = ("%d steps x %d images = one rank's batch of the same workload (numpy gaussian_noise per-image loop + torch fp32 "
              "ResNet-50), all %d host threads; one process, as the reference runs AddNoise" % (args.steps, per_step, threads))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(world),
            "reference_note": "reference CPU path = oracle port (the reference package itself cannot be imported: skimage / wand / "
                              "easydict absent); rank 0 alone runs it, on a bounded sample of %d images per step" % per_step,
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
DTYPE = {"split": "f16x3 (every fp32 value = fp16 hi + fp16 lo plane, three fp16 MMAs per product, fp32 accumulation in TMEM: "
                  "fp32-faithful, logits within 1e-3 of the reference at realistic magnitude -- tests/test_calibrated_gpu.py)",
         "f16": "f16 (ONE fp16 plane, one MMA per product, fp32 accumulation: TF32-class, 2.6e-2 .. 4.7e-2 max logit error at "
                "realistic magnitude, NOT within the 1e-3 north-star tolerance)"}
PASSES = {"split": 3, "f16": 16}


def small_kernels_ms(model, pipe, reps=20):
    """CUDA-event time of the non-tensor-core kernels of a step besides the corruption: global average pool + counters."""
    import torch
    from robustart_b200 import ops
    n = pipe.static_in.shape[0]
    f16 = getattr(model, "f16", False)
    c = model.feat if hasattr(model, "feat") else model.fc_w.shape[-1]
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


def gemm_inventory(model, pipe, precision):
    """Algorithmic FLOPs and activation bytes of every tensor-core launch of one forward, from the shapes the launches
    are made with (one eager forward with recording wrappers; nothing is timed here)."""
    import torch
    from robustart_b200 import ops
    recs = []
    orig = (ops.conv2d_nhwc, ops.linear, ops.stem_conv7x7_u8, ops.stem_pool_u8)

    def conv(x, wgt, *a, **k):
        _, n, h, w, cin = x.shape
        _, cout, kh, kw, _ = wgt.shape
        st, pd = k.get("stride", 1), k.get("pad", 0)
        ho, wo = (h + 2 * pd - kh) // st + 1, (w + 2 * pd - kw) // st + 1
        has_res = (a[2] if len(a) > 2 else k.get("res")) is not None
        bpe = 2.0 if x.shape[0] == 1 else 4.0            # one fp16 plane or two fp16 planes per element
        recs.append((2.0 * n * ho * wo * cout * kh * kw * cin, bpe * (n * h * w * cin + n * ho * wo * cout * (2 if has_res else 1))))
        return orig[0](x, wgt, *a, **k)

    def lin(x, wgt, *a, **k):
        kk = x.shape[-1]
        rows = x[0].numel() // kk
        recs.append((2.0 * rows * kk * wgt.shape[1], (2.0 if x.shape[0] == 1 else 4.0) * (rows * kk + rows * wgt.shape[1])))
        return orig[1](x, wgt, *a, **k)

    def stem(img, *a, **k):
        n, hh, ww = img.shape[0], img.shape[1] // 2, img.shape[2] // 2
        recs.append((2.0 * n * hh * ww * 64 * 147, img.numel() + (2.0 if a[0].shape[0] == 1 else 4.0) * n * hh * ww * 64))
        return orig[2](img, *a, **k)

    def stem_pool(img, *a, **k):       # one-launch stem (conv1 + bn + relu + maxpool on the tensor core, fp16 mode)
        n = img.shape[0]
        recs.append((2.0 * n * (img.shape[1] // 2) * (img.shape[2] // 2) * 64 * 147,
                     img.numel() + 2.0 * n * (img.shape[1] // 4) * (img.shape[2] // 4) * 64))
        return orig[3](img, *a, **k)

    ops.conv2d_nhwc, ops.linear, ops.stem_conv7x7_u8, ops.stem_pool_u8 = conv, lin, stem, stem_pool
    try:
        model.forward(pipe.static_in)
        torch.cuda.synchronize()
    finally:
        ops.conv2d_nhwc, ops.linear, ops.stem_conv7x7_u8, ops.stem_pool_u8 = orig
    return recs


def gemm_roofline(model, pipe, pk, precision, step_ms, other_ms):
    """The dominant kernel set: all tcgen05 launches of one forward (implicit-GEMM convolutions, stem, fc).  Time = the
    step's device time INSIDE the timed region minus the step's other kernels (corruption, avgpool, counters -- each
    timed alone); achieved = algorithmic FLOPs / that time."""
    recs = gemm_inventory(model, pipe, precision)
    flops = sum(f for f, _ in recs)
    mmas = 1.0 if precision == "f16" else 3.0
    one_pass_peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])     # fp16 and bf16 MMAs run at the same rate
    peak = one_pass_peak / mmas
    gemm_ms = step_ms - other_ms
    achieved = flops / gemm_ms / 1e9                                       # FLOP / ms / 1e9 = TFLOP/s
    layer_ms = sum(max(b / (pk["hbm_gbs"] * 1e9), mmas * f / (one_pass_peak * 1e12)) for f, b in recs) * 1e3
    traffic = None
    for name in ("r2_gemm_traffic_pairs.json" if precision == "split" else "", "r2_gemm_traffic_%s.json" % precision,
                 "r1_gemm_traffic_%s.json" % {"split": "bf16x3"}.get(precision, precision)):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):      # dram__bytes_read+write summed over the same launches, from one ncu capture (profiles/)
            traffic = json.load(open(tp)).get("dram_bytes_per_forward")
            traffic_src = name
            break
    return {"bound": "tensor", "kernel": "gemm_kernel<BN> (+ stem) -- tcgen05 implicit GEMMs, all %d launches of one forward" % len(recs),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_note": ("bytes per forward over the same launches, ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                                                 "profiles/%s" % traffic_src) if traffic else None,
            "peak_source": pk["_source"] + ": sustained dense 16-bit tensor throughput (kernel timed inside a long step)" +
                           (" / 3: this precision issues three MMAs per fp32 product, so the tensor pipe can deliver a third of its "
                            "one-pass rate in algorithmic FLOPs (SURVEY 8d: peak of the precision actually used)" if mmas == 3 else ""),
            "one_pass_peak": one_pass_peak, "frac_of_one_pass_peak": achieved / one_pass_peak,
            "time_source": "ms_per_step (CUDA events around the timed loop) minus the corruption / avgpool / counter kernels timed alone "
                           "(%.3f ms): %.3f ms of tensor-core launches per step" % (other_ms, gemm_ms),
            "gemm_ms_per_step": gemm_ms, "algorithmic_gflop_per_step": flops / 1e9,
            "layer_roofline_ms_per_step": layer_ms, "frac_of_layer_roofline": layer_ms / gemm_ms,
            "layer_roofline_note": "sum over the launches of max(algorithmic activation bytes / HBM peak, issued MMA FLOPs / one-pass tensor "
                                   "peak): many 1x1 layers of ResNet-50 are bound by activation traffic, not by the tensor pipe"}


def corruption_roofline(pipe, inputs, pk):
    """Kernel-only time of the gaussian_noise launch at the bench batch (256 images = 77 MB of traffic, ~12 us at
    HBM speed -- less than a Python launch), so the launches are captured into a CUDA graph (one per rotating
    input batch) and the graph is replayed: CUDA-event time / launches = device time per launch."""
    import torch
    from robustart_b200 import ops
    outs = [torch.empty_like(inputs[0]) for _ in range(len(inputs))]
    for i in range(5):               # every severity once, eagerly: the kernel's quantile table is uploaded at its first use (not capturable)
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
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    reps, groups, times = 10, 7, []
    for _ in range(groups):          # median of 7 groups: a 3 ms measurement right after the long loops sees clock ramps
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e) * 1e-3 / (reps * per_graph))
    t = sorted(times)[groups // 2]
    alg = 2.0 * BATCH * H * W * 3
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r2_ncu_corruptions.json")
    if os.path.exists(tp):
        for row in json.load(open(tp)).get("kernels", []):
            if row.get("corruption") == "gaussian_noise" and row.get("dram_bytes") is not None:
                traffic = row["dram_bytes"]
    return {"bound": "hbm", "kernel": "normal_noise_strata_kernel (gaussian_noise, u8 NHWC -> u8 NHWC, device Philox + quantile table)", "achieved": alg / t / 1e9,
            "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": alg / t / 1e9 / pk["hbm_gbs"], "traffic": traffic,
            "us_per_launch": t * 1e6, "algorithmic_bytes_per_launch": alg, "images_per_s_kernel_only": BATCH / t,
            "us_per_launch_groups": [round(x * 1e6, 2) for x in times],
            "traffic_note": "ncu dram bytes of a warm launch on ONE repeated batch (profiles/r2_ncu_corruptions.json): the 126 MB L2 holds the previous output, so less than the algorithmic bytes reach DRAM there; the timed launches rotate over 8 batches",
            "timing": "CUDA graph of %d launches over %d rotating input batches (> L2), median of %d groups of %d replays" % (per_graph, len(inputs), groups, reps)}


def timed_value(pipe, inputs, labels, steps, warmup, barrier):
    """`value`: K device-resident steps + the evaluation's ONE collective (counter all-reduce), all between the two events."""
    import torch
    for i in range(warmup):
        pipe.step_device(inputs[i % len(inputs)], labels, "gaussian_noise", 1 + i % 5)
    pipe.reset()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        pipe.step_device(inputs[i % len(inputs)], labels, "gaussian_noise", 1 + i % 5)
    pipe.allreduce_counters()                      # NCCL all-reduce of int64[3] over NVLink, on the compute stream
    e.record()
    barrier()
    return s.elapsed_time(e)


def run_ours(args):
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
    precision = {"bf16x3": "split"}.get(args.precision, args.precision)

    sd = resnet50_weights()
    py_model = nets.build_model("resnet50", sd, device=dev, passes=PASSES[precision])      # Python layer sequencing (attack side reports, shapes)
    if args.sequencing == "handle":
        from robustart_b200.handle import ModelHandle
        model = ModelHandle("resnet50", sd, dev, PASSES[precision])                       # C++ layer sequencing behind the C-ABI (b200r_model_*)
    else:
        model = py_model
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

    if world > 1:          # NCCL communicator bring-up outside the timed region
        pipe.allreduce_counters()
        pipe.reset()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # the corruption kernel alone (its roofline entry) is timed BEFORE the long loops: it is a 17 us issue-bound kernel whose time
    # follows the SM clock, and right after seconds of tensor-core load the power cap still holds the clock at ~1.5 GHz
    roof_c = None
    if rank == 0:
        try:
            roof_c = corruption_roofline(pipe, inputs, pk)
        except Exception as ex:               # an auxiliary measurement must never take the headline line down
            roof_c = {"error": "%s: %s" % (type(ex).__name__, ex), "us_per_launch": 20.0}
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ---------------------------------------------------
    dev_ms = timed_value(pipe, inputs, labels, args.steps, args.warmup, barrier)
    counts = pipe.counters.tolist()

    # ---- end-to-end through host buffers ("e2e") ----------------------------------------------
    pipe.reset()
    for i in range(max(3, args.warmup // 2)):
        pipe.step_host_pipelined(h_inputs[i % len(h_inputs)], h_labels, "gaussian_noise", 1 + i % 5)
    pipe.finish()
    pipe.reset()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for i in range(args.steps):
        pipe.step_host_pipelined(h_inputs[i % len(h_inputs)], h_labels, "gaussian_noise", 1 + i % 5)
    pipe.finish()                          # the last step's counters are on the host
    pipe.allreduce_counters()              # the collective, then the job's result on the host
    host_counters = pipe.counters.cpu()
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
        try:
            other_ms = roof_c["us_per_launch"] * 1e-3 + small_kernels_ms(model, pipe)
            roof = gemm_roofline(py_model, pipe, pk, precision, dev_ms / args.steps, other_ms)
        except Exception as ex:               # never lose the headline line to an auxiliary measurement
            roof = {"error": "%s: %s" % (type(ex).__name__, ex)}
        clocks = sampler.stop()
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": DTYPE[precision], "data": "synthetic",
                "config": make_config(world),
                "impl_detail": {"forward": "CUDA graph replay of %s" % ("b200r_model_forward_u8 (C++ layer sequencing behind the C-ABI, csrc/model_handle.cu)"
                                                                        if args.sequencing == "handle" else "nets.ResNet.forward (Python layer sequencing)"),
                                "counters": [int(c) for c in counts],
                                "collective": "one all-reduce of int64[3] inside the timed region" if world > 1 else "none at N=1"},
                "clocks": clocks,
                "e2e": {"value": e2e_v, "unit": "images/s", "h2d_bytes_per_step": BATCH * H * W * 3 + BATCH * 8,
                        "d2h_bytes_per_step": 24, "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall_ms / args.steps,
                        "api": "CorruptEvalPipeline.step_host_pipelined + finish (pinned host uint8 batch -> H2D on a copy stream -> device step -> "
                               "counters D2H every step; the host waits for step i-1's counters while step i runs), then the counter all-reduce",
                        "counters_on_host": [int(v) for v in host_counters.tolist()]},
                "gpu_launches": pipe.launches_per_step * args.steps,
                "gpu_launches_per_step": pipe.launches_per_step,
                "roofline": roof, "roofline_corruption": roof_c}
        def guarded(fn, *a):
            try:
                return fn(*a)
            except Exception as ex:                       # auxiliary reports must never take the headline line down
                return {"error": "%s: %s" % (type(ex).__name__, ex)}
        if world == 1 and not args.no_side:
            line["modes"] = guarded(other_mode_report, precision, value, sd, dev, inputs, labels, barrier, pk)
        if world == 1 and not args.no_cpu_baseline:
            r = guarded(cpu_baseline_and_match, model, dev)
            line["cpu_baseline"], line["top1_match"] = r if isinstance(r, tuple) else (r, r)
        if world == 1 and not args.no_pgd:
            line["pgd_loop"] = guarded(pgd_loop_report, py_model, dev, pk, precision)
        if world == 1 and not args.no_side:
            for key, fn in (("configs2_vit_pgd", side_vit_pgd), ("configs3_sweep", side_sweep), ("configs4_mixer_aa", side_mixer_aa)):
                try:
                    line[key] = fn(dev, pk)
                except Exception as ex:                   # a side report must never take the headline line down
                    line[key] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_mode_report(precision, value, sd, dev, inputs, labels, barrier, pk):
    """Both precision modes' device-resident img/s: the default is the one that holds the 1e-3 tolerance."""
    from robustart_b200 import nets
    from robustart_b200.evalpipe import CorruptEvalPipeline
    other = "f16" if precision == "split" else "split"
    m2 = nets.build_model("resnet50", sd, device=dev, passes=PASSES[other])
    p2 = CorruptEvalPipeline(m2, BATCH, H, W, seed=99)
    ms = timed_value(p2, inputs, labels, 20, 5, barrier)
    return {precision: {"images_per_s": value, "dtype": DTYPE[precision], "default": True},
            other: {"images_per_s": BATCH * 20 / (ms * 1e-3), "dtype": DTYPE[other], "default": False}}


def cpu_baseline_and_match(model, dev, n=256):
    """(cpu_baseline, top1_match).  One bounded sample serves both: 256 structurally different images through the CPU port
    (numpy gaussian_noise loop + torch fp32 ResNet-50), timed; the SAME images with the SAME normal draws (ext_noise) through
    the GPU path; predictions and logits compared."""
    import numpy as np
    import torch
    from robustart_b200 import ops
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import diverse_images, ext_from_log
    threads = os.cpu_count() or 1
    cpu_reference_pass(32, threads)                                   # warm the host path
    clean = diverse_images(n, seed=5)
    sev = 3
    r = cpu_reference_pass(n, threads, seed=sev - 1, images=clean.copy(), keep_logits=True)
    ext = np.concatenate([np.asarray(a, dtype=np.float64).ravel() for _, a in r["draws"].log]).astype(np.float32)
    got_img = ops.corrupt_u8(torch.from_numpy(clean).to(dev), "gaussian_noise", sev, ext_noise=torch.from_numpy(ext).to(dev))
    logits = model.forward(got_img).float().cpu()
    ref = r["logits"]
    d_img = np.abs(got_img.cpu().numpy().astype(np.int16) - r["corrupted"].astype(np.int16))
    # logits on IDENTICAL inputs (the CPU-corrupted batch through the GPU model): isolates the model's arithmetic
    same_in = model.forward(torch.from_numpy(r["corrupted"]).to(dev)).float().cpu()
    top5_ref = ref.topk(5, dim=1).indices
    cpu = {"value": n / r["total_s"], "unit": "images/s", "cores": threads, "kind": "port",
           "sample": "%d images of the same workload: oracle numpy gaussian_noise per-image loop (%.0f img/s alone) + torch fp32 ResNet-50 "
                     "forward (%.0f img/s alone)" % (n, n / r["noise_s"], n / r["model_s"])}
    match = {"images": n, "severity": sev, "shared_draws": "the CPU port's numpy normals passed to the GPU kernel as ext_noise",
             "top1_agreement": float((logits.argmax(1) == ref.argmax(1)).float().mean()),
             "top1_in_reference_top5": float((logits.argmax(1, keepdim=True) == top5_ref).any(1).float().mean()),
             "distinct_reference_top1_classes": int(ref.argmax(1).unique().numel()),
             "max_abs_dlogit": float((logits - ref).abs().max()),
             "max_abs_dlogit_identical_input": float((same_in - ref).abs().max()),
             "reference_logit_absmax": float(ref.abs().max()), "reference_logit_std": float(ref.std()),
             "corrupted_bytes_differing": float((d_img > 0).mean()), "corrupted_bytes_max_diff": int(d_img.max())}
    return cpu, match


def pgd_loop_report(model, dev, pk, precision, n=128, steps=10, reps=2):
    """The second half of the north star, reported beside the headline metric (not part of `value`): the PGD-Linf
    10-step eval loop on ResNet-50, forward + input gradient on the sm_100a kernels, then one
    forward of the adversarial batch + counters.  Algorithmic FLOPs = (2k+1) * 8.18 GFLOP / image."""
    return _pgd_report("resnet50", model, dev, pk, precision, n, steps, reps, RESNET50_GFLOP)


def _pgd_report(arch, model, dev, pk, precision, n, steps, reps, fwd_gflop):
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
    tf = (2 * steps + 1) * fwd_gflop * n / ms
    mmas = 1.0 if precision == "f16" else 3.0
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) / mmas
    return {"workload": "%s, pgd_linf eps 4/255, %d steps, batch %d, float32 NCHW images resident in HBM" % (arch, steps, n),
            "images_per_s": n / ms * 1e3, "ms_per_batch": ms, "algorithmic_tflops": tf, "peak_tflops": peak, "frac": tf / peak,
            "source_model": "native input-gradient pass (tcgen05 GEMMs on transposed weights, %s), no autograd" % precision,
            "note": "gradients run loss-scaled (x4096), unscaled when the image gradient is written; peak = sustained 16-bit tensor "
                    "throughput / MMAs per product (%d)" % int(mmas)}


def side_vit_pgd(dev, pk):
    """BASELINE configs[2]: ViT-B/16 + PGD-Linf eps 4/255, 10 steps, batch 128 (one GPU)."""
    from robustart_b200 import nets
    model = nets.build_model("vit_b16_224", device=dev, passes=3)
    return _pgd_report("vit_b16_224", model, dev, pk, "split", 128, 10, 1, 35.1)


def side_sweep(dev, pk):
    """BASELINE configs[3] on ONE GPU: MobileNetV2 / EfficientNet-B0 + the full ImageNet-C 15 x 5 sweep, batch 256 per cell."""
    import torch
    from robustart_b200 import nets, ops
    names = ["gaussian_noise", "shot_noise", "impulse_noise", "defocus_blur", "glass_blur", "motion_blur", "zoom_blur", "snow", "frost",
             "fog", "brightness", "contrast", "elastic_transform", "pixelate", "jpeg_compression"]
    g = torch.Generator(device=dev).manual_seed(3)
    imgs = torch.randint(0, 256, (BATCH, H, W, 3), dtype=torch.uint8, device=dev, generator=g)
    labels = torch.randint(0, 1000, (BATCH,), device=dev, generator=g)
    work = torch.empty_like(imgs)
    out = {"workload": "15 corruptions x 5 severities = 75 cells x %d images per pass, corruption kernel -> graphed forward -> counters" % BATCH}
    for arch in ("mobilenet_v2", "efficientnet_b0"):
        model = nets.build_model(arch, device=dev, passes=3)
        run = model.graphed(work)
        counters = torch.zeros((75, 3), dtype=torch.int64, device=dev)

        def sweep():
            for ci in range(75):
                ops.corrupt_u8(imgs, names[ci // 5], 1 + ci % 5, seed=7, out=run.static_in)
                ops.topk_count_(counters[ci], run(run.static_in, copy_in=False), labels)

        sweep()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        sweep()
        e.record()
        torch.cuda.synchronize()
        out[arch] = {"corrupted_images_per_s": 75 * BATCH / (s.elapsed_time(e) * 1e-3), "ms_per_sweep": s.elapsed_time(e)}
    return out


def side_mixer_aa(dev, pk):
    """BASELINE configs[4] on ONE GPU: MLP-Mixer-B/16 + AutoAttack-Linf eps 4/255, batch 64.  (a) the standard pipeline as the plugin
    runs it (apgd-ce -> apgd-t -> fab-t -> square on the shrinking robust set; with synthetic weights APGD-CE already fools every sample,
    so the later stages see an empty set); (b) every stage alone on the full batch with a reduced budget (3 target classes instead of 9,
    500 Square queries instead of 5000 -- the stages' cost is linear in both), so that each stage's device path is timed."""
    import torch
    from robustart_b200 import attacks, autoattack, nets
    model = nets.build_model("mixer_b16_224", device=dev, passes=3)
    src = attacks.NativeModel(model)
    g = torch.Generator(device=dev).manual_seed(4)
    n = 64
    x = torch.rand(n, 3, 224, 224, device=dev, generator=g)
    y = model.forward(x).argmax(1)                       # labels = clean predictions: every sample starts robust
    torch.cuda.synchronize()
    # one untimed pass first: the first gradient call builds the transposed weight planes and sizes the allocator's pools
    autoattack.AutoAttack(src, norm="Linf", eps=4 / 255, seed=0, verbose=False, version="standard").run_standard_evaluation(x, y, bs=n)
    torch.cuda.synchronize()
    aa = autoattack.AutoAttack(src, norm="Linf", eps=4 / 255, seed=0, verbose=False, version="standard")
    t0 = time.perf_counter()
    adv = aa.run_standard_evaluation(x, y, bs=n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"workload": "mixer_b16_224, AutoAttack-Linf standard (apgd-ce, apgd-t x9, fab-t x9, square 5000), eps 4/255, batch %d, "
                       "labels = clean predictions" % n,
           "images_per_s": n / dt, "seconds_per_batch": dt, "robust_accuracy_after_each_stage": aa.history,
           "linf": float((adv - x).abs().max()), "source_model": "native input-gradient pass (token_backward.cu + dgrad GEMMs)",
           "stages_alone": {}}
    for stage, kw in (("apgd-ce", {}), ("apgd-t", {"n_target_classes": 3}), ("fab-t", {"n_target_classes": 3}), ("square", {"n_queries": 500})):
        one = autoattack.AutoAttack(src, norm="Linf", eps=4 / 255, seed=0, verbose=False, version="standard", attacks_to_run=[stage], **kw)
        f0, b0 = one.m.forwards, one.m.backwards
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        one.run_standard_evaluation(x, y, bs=n)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["stages_alone"][stage] = {"seconds": dt, "budget": kw or "standard (100 iterations)", "forward_images": one.m.forwards - f0,
                                      "backward_images": one.m.backwards - b0, "robust_after": one.history[-1][1]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-pgd", action="store_true", help="skip the PGD-loop side report")
    ap.add_argument("--no-side", action="store_true", help="skip the configs[2] / [3] / [4] side reports and the second precision mode")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sequencing", default="handle", choices=["handle", "python"],
                    help="who issues the forward's launches: the C++ model handle behind the C-ABI (default) or robustart_b200/nets.py")
    ap.add_argument("--precision", default="split", choices=["split", "f16", "bf16x3"],
                    help="split: fp16 hi/lo planes, three MMAs per product, fp32-faithful (default; `bf16x3` is the old name); "
                         "f16: one fp16 plane, one MMA per product, TF32-class")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
