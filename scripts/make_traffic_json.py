"""profiles/ncu_traffic.json from `ncu --set full` captures: per bench-profile kernel name, dram__bytes_read.sum + dram__bytes_write.sum of ONE
launch of its dominant layer shape, next to the algorithmic bytes of that launch (bench.py reports both in `roofline.traffic_detail`).

    python scripts/make_traffic_json.py          # reads the captures listed in CAPTURES (gpurun_out/*.ncu-rep), rewrites profiles/ncu_traffic.json
"""
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# bench kernel name -> (capture, substring of the ncu kernel name, launch index among the matches, layer description, algorithmic bytes)
PX = 1088 * 1920
CAPTURES = {
    "conv2d_tc_3xf16": ("gpurun_out/r02d_ncu_3xf16_gru.ncu-rep", "conv2d_tc_kernel", 0, "1x5 384->128 @2x136x240 (SepConvGRU gate), 3xF16 split kernel on CTA pairs",
                        2 * 136 * 240 * (384 + 128) * 4.0),
    "hyponet_fused3": ("gpurun_out/r02k_ncu_hyponet.ncu-rep", "hyponet_fused3", 0, "fused 5-layer HypoNet, fp32-class, 1088x1920 pixels", PX * (32 + 3 + 2) * 4.0),
    "hyponet_fused": ("gpurun_out/r02k_ncu_hyponet.ncu-rep", "hyponet_fused_kernel", 0, "fused 5-layer HypoNet, TF32/half operands, 1088x1920 pixels", PX * (32 + 3 + 2) * 4.0),
    "conv2d_tc_f16": ("gpurun_out/r02k_ncu_f16_trunk.ncu-rep", "conv2d_tc_kernel", 0, "3x3 256->256 @1x1088x1920, fp16 storage, kind::f16 on CTA pairs (final-decoder trunk)",
                      PX * (256 + 256) * 2.0),
    "conv2d_halo_tf32": ("gpurun_out/r02k_ncu_halo.ncu-rep", "conv3x3_halo", 0, "3x3 32->32 @2x1088x1920, halo-reuse kernel (direct-store epilogue)", 2 * PX * (32 + 32) * 4.0),
    "softsplat_fused": ("gpurun_out/r02i_ncu_splat.ncu-rep", "softsplat_tile", 0, "one-pass tile splat 16+1 ch @1088x1920 (off by default)", PX * 140.0),
    "softsplat_accumulate": ("gpurun_out/r02a_hbm_kernels.ncu-rep", "softsplat_acc", 0, "forward splat 16+1 ch @1088x1920 (accumulate pass)", PX * 140.0),
    "softsplat_normalize": ("gpurun_out/r02a_hbm_kernels.ncu-rep", "SplatNorm", 0, "zeroeps normalisation @1088x1920", PX * (20 + 16) * 4.0),
    "backwarp": ("gpurun_out/r02a_hbm_kernels.ncu-rep", "Backwarp", 0, "backward warp 64 ch @1088x1920", PX * (2 * 64 + 2) * 4.0),
    "resize_bilinear": ("gpurun_out/r02a_hbm_kernels.ncu-rep", "ResizeK", 0, "bilinear x4 128 ch 272x480 -> 1088x1920", 272 * 480 * 128 * 4.0 * 17),
    "instnorm_partial": ("gpurun_out/r02a_hbm_kernels.ncu-rep", "InPartial", 0, "instance-norm statistics 64 ch @2x544x960", 2 * 544 * 960 * 64 * 4.0),
    "instnorm_apply": ("gpurun_out/r02a_hbm_kernels.ncu-rep", "InApply", 0, "instance-norm apply + ReLU 64 ch @2x544x960", 2 * 544 * 960 * 64 * 8.0),
    "convex_upsample": ("gpurun_out/r02a_hbm_kernels.ncu-rep", "ConvexUp", 0, "convex x8 upsample 2x136x240 -> 2x1088x1920", (2 * 136 * 240 * 578 + 2 * 2 * PX) * 4.0),
    "corr_lookup": ("gpurun_out/r02k_ncu_lookup.ncu-rep", "corr_lookup_warp", 0, "4-level 9x9 correlation lookup from the volume pyramid, 136x240 source pixels (warp per pixel and level)",
                    136 * 240 * (324 + 400) * 4.0),
    "corr_lookup_direct": ("gpurun_out/r02k_ncu_lookup.ncu-rep", "corr_lookup_direct", 0, "volume-free 4-level 9x9 lookup (100 dots of 256 ch per pixel and level), 136x240 source pixels",
                           136 * 240 * ((256 + 324) * 4.0 + 256 * 2.0 * (1 + 1 / 4 + 1 / 16 + 1 / 64))),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def rows_of(path):
    out = subprocess.run(["ncu", "-i", os.path.join(ROOT, path), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    cache, res = {}, {}
    for key, (path, pat, idx, desc, alg) in CAPTURES.items():
        if not os.path.exists(os.path.join(ROOT, path)):
            continue
        if path not in cache:
            cache[path] = rows_of(path)
        hdr, units, rows = cache[path]
        ki = hdr.index("Kernel Name")
        m = [r for r in rows if pat in r[ki]]
        if len(m) <= idx:
            continue
        r = m[idx]
        val = lambda name, table: float(r[hdr.index(name)]) * table[units[hdr.index(name)]]
        rd, wr, ms = val("dram__bytes_read.sum", UNIT), val("dram__bytes_write.sum", UNIT), val("gpu__time_duration.sum", TIME)
        e = {"layer": desc, "bytes": rd + wr, "dram_read": rd, "dram_write": wr, "algorithmic_bytes": alg, "ratio_to_algorithmic": (rd + wr) / alg,
             "ncu_duration_ms": ms, "capture": os.path.basename(path).replace(".ncu-rep", "")}
        tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
        if tp in hdr and r[hdr.index(tp)] not in ("", "n/a"):
            try:
                e["tensor_pipe_active_pct"] = float(r[hdr.index(tp)])
            except ValueError:
                pass
        res[key] = e
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res.items():
        print("%-22s %8.1f MB dram (%.2fx algorithmic) %.3f ms  %s" % (k, v["bytes"] / 1e6, v["ratio_to_algorithmic"], v["ncu_duration_ms"], v["layer"]))


if __name__ == "__main__":
    main()
