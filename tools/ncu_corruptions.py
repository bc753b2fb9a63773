"""profiles/rN_ncu_corruptions.json from one ncu capture of all 19 ImageNet-C corruption launches (read here, no GPU):

    NREP=2 ncu --section SpeedOfLight --section WarpStateStats --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis \
        --section ComputeWorkloadAnalysis --section SchedulerStats --clock-control none -o gpurun_out/rN_corruptions \
        python tools/ncu_targets.py <the 19 names>                                   (on the GPU box)
    python tools/ncu_corruptions.py gpurun_out/rN_corruptions.ncu-rep profiles/rN_ncu_corruptions.json

Per kernel (last captured launch = warm): duration, DRAM bytes, DRAM / L1 / issue / pipe utilisation, occupancy, the top stall
reasons, and the limiter they point to.  Batch = 256 images of 224 x 224 x 3 (algorithmic bytes 77.07 MB per corruption)."""
import csv
import io
import json
import subprocess
import sys

KERNEL_TO_CORRUPTION = {
    "normal_noise_strata_kernel<0>": "gaussian_noise", "normal_noise_strata_kernel<1>": "speckle_noise",
    "normal_noise_strata_kernel<0, 1024, 1, 1>": "gaussian_noise", "normal_noise_strata_kernel<1, 1024, 1, 1>": "speckle_noise",
    "normal_noise_icdf_kernel<0>": "gaussian_noise", "normal_noise_icdf_kernel<1>": "speckle_noise",
    "normal_noise_rng_kernel<0>": "gaussian_noise", "normal_noise_rng_kernel<1>": "speckle_noise",
    "shot_kernel<0>": "shot_noise", "impulse_rng_kernel": "impulse_noise", "defocus_kernel": "defocus_blur",
    "glass_shuffle_kernel": "glass_blur", "gauss_blur_kernel": "gaussian_blur (+ glass_blur's two blurs)", "motion_blur_kernel": "motion_blur",
    "zoom_blur_kernel": "zoom_blur", "snow_kernel": "snow", "frost_kernel": "frost", "fog_plasma_kernel": "fog (plasma generator)",
    "fog_blend_kernel": "fog (blend)", "hsv_kernel": "brightness / saturate", "channel_sum_kernel": "contrast (channel sums)",
    "contrast_apply_kernel": "contrast (apply)", "elastic_warp_kernel": "elastic_transform (affine warp)",
    "elastic_field_kernel": "elastic_transform (random fields)", "elastic_matmul_kernel": "elastic_transform (Gaussian as two dense products)",
    "elastic_gather_kernel": "elastic_transform (map_coordinates)", "pixelate_kernel": "pixelate", "jpeg_kernel": "jpeg_compression",
    "spatter_layer_kernel": "spatter (liquid layer)", "plane_blur_kernel": "spatter (plane blurs)", "spatter_water_kernel": "spatter (water, sev 1-3)",
    "spatter_mud_kernel": "spatter (mud, sev 4-5)",
}
SCALE = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6,
         "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
M = {"us": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
     "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread", "grid": "launch__grid_size",
     "block": "launch__block_size"}
STALLS = ["long_scoreboard", "short_scoreboard", "math_pipe_throttle", "mio_throttle", "wait", "not_selected", "barrier", "lg_throttle",
          "dispatch_stall", "no_instruction", "branch_resolving", "tex_throttle", "membar", "sleeping", "drain", "imc_miss"]


def limiter(e):
    if e["dram_pct"] >= 60:
        return "HBM bandwidth"
    if e["l1tex_pct"] >= 70:
        return "L1 / shared-memory bandwidth (gathers, staging traffic)"
    if e["warps_pct"] < 20:
        return "latency of a serial chain (few resident warps by construction)"
    top = e["stalls"][0][0] if e["stalls"] else ""
    if e["issue_pct"] >= 75 or top in ("math_pipe_throttle", "not_selected", "dispatch_stall"):
        return "instruction issue / math pipes (top stall: %s)" % top
    if top in ("long_scoreboard", "lg_throttle"):
        return "memory latency (top stall: %s)" % top
    if top in ("barrier", "short_scoreboard", "mio_throttle"):
        return "shared memory / barriers (top stall: %s)" % top
    return "mixed (top stall: %s)" % top


def main(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    last = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()
        if name.startswith("at::"):
            continue
        last[name] = d

    def val(d, key):
        v = d.get(key, "")
        if v == "":
            return 0.0
        return float(v.replace(",", "")) * SCALE.get(unit.get(key, ""), 1.0)
    out = []
    for name, d in last.items():
        e = {"kernel": name, "corruption": KERNEL_TO_CORRUPTION.get(name, "?")}
        for k, m in M.items():
            e[k] = round(val(d, m), 3)
        rdwr = (e.pop("rd") + e.pop("wr")) * 1e6
        if rdwr == 0 and d.get("dram__bytes.sum.per_second", "") != "":
            # section captures hold the rate, not the byte counters: bytes = rate x duration
            per_s = float(d["dram__bytes.sum.per_second"].replace(",", "")) * {"Gbyte/s": 1e9, "Tbyte/s": 1e12, "Mbyte/s": 1e6, "Kbyte/s": 1e3, "byte/s": 1.0}.get(unit.get("dram__bytes.sum.per_second", ""), 1.0)
            rdwr = per_s * e["us"] * 1e-6
        e["dram_bytes"] = round(rdwr) if rdwr > 0 else None
        st = []
        for s in STALLS:
            key = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if d.get(key, "") != "":
                st.append((s, round(float(d[key].replace(",", "")), 2)))
        e["stalls"] = sorted(st, key=lambda t: -t[1])[:3]
        e["limiter"] = limiter(e)
        out.append(e)
        print("%-34s %-40s %8.1f us  dram %5.1f%% l1 %5.1f%% issue %5.1f%% warps %5.1f%%  %s" % (
            name[:34], e["corruption"][:40], e["us"], e["dram_pct"], e["l1tex_pct"], e["issue_pct"], e["warps_pct"], e["limiter"]))
    json.dump({"batch": "256 x 224 x 224 x 3 uint8 (algorithmic bytes 77 070 336 per corruption)",
               "capture": "ncu (sections SpeedOfLight, WarpStateStats, Occupancy, LaunchStats, MemoryWorkloadAnalysis, ComputeWorkloadAnalysis, "
                          "SchedulerStats; --clock-control none), last (warm) launch of each kernel; times under the profiler, cold L2",
               "kernels": out}, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
