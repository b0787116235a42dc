#!/usr/bin/env python
"""bench.py — interpolated frames/s of the GIMM-VFI-R per-pair path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one GIMMVFI_R.forward over one batch of synthetic frame pairs.  Workload at
every N: BASELINE.json configs[1] — one 1920x1080 pair per GPU (caller-padded to
1088x1920), t=0.5, GIMM-VFI-R, random-init weights (no checkpoints offline), all
reference outputs produced.  N>1: one process per GPU (torchrun), pairs sharded with no
data-path collective, ONE all-gather of the output frames per step (weak scaling).

Prints one JSON line (rank 0).  `value` = device-timed frames/s with inputs resident in
HBM; `e2e` = the same through the public API with pinned-host inputs, H2D + D2H inside the
timed region; `roofline` = the dominant kernel timed live with CUDA events on the
launching stream; `cpu_baseline` = the CPU oracle port on the box's host cores (bounded
sample).  `--impl reference` times that CPU path only (the reference is Python/PyTorch:
its own path on host cores is the oracle port, bit-identical to it here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "interpolated frames/sec @1080p t=0.5"
UNIT = "frames/s"
H_PAD, W_PAD = 1088, 1920          # InputPadder(1080x1920, 32)  src/utils/utils.py:156-185
SAMPLE_H, SAMPLE_W = 256, 448      # bounded CPU sample (BASELINE config 1 size)
if os.environ.get("GIMMVFI_CPU_SAMPLE"):   # tests shrink the sample (e.g. "128x160"); the scaling to 1080p is by pixel count either way
    SAMPLE_H, SAMPLE_W = (int(v) for v in os.environ["GIMMVFI_CPU_SAMPLE"].lower().split("x"))


def flops_per_frame(P, T=1, P_full=None):
    """algorithmic FLOPs (2*MAC, conv+matmul) of GIMM-VFI-R for one pair."""
    P_full = P if P_full is None else P_full
    return 4.981e6 * P + 0.375 * float(P) ** 2 + T * (13.770e6 * P + 0.0212e6 * P_full)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), which="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, which="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max((float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_threads():
    """Threads for the CPU arm: all host cores up to 32 (the reference's small convolutions stop scaling —
    and oversubscribe oneDNN — beyond that on the 128-core GPU boxes)."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("GIMMVFI_CPU_THREADS", "32"))))


def cpu_reference_fps(steps, warmup):
    """The reference's own PyTorch path on the host cores, via the oracle port (bit-identical
    to the reference in the build container, tests/golden/manifest.json).  Bounded sample:
    one 256x448 pair per step; frames/s scaled by the pixel ratio to the 1088x1920 workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gimmvfi_r_oracle as O
    from gimmvfi_b200.synth import synth_batch
    from gimmvfi_b200.weights import random_state_dict

    torch.set_num_threads(cpu_threads())
    sd = random_state_dict(0)
    xs = synth_batch(1, SAMPLE_H, SAMPLE_W, seed=6)
    coord = [(O.sample_coord_input(1, (SAMPLE_H, SAMPLE_W), [0.5]), None)]
    t = [0.5 * torch.ones(1)]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.gimmvfi_r_forward(sd, xs, coord, t)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    ratio = (SAMPLE_H * SAMPLE_W) / float(H_PAD * W_PAD)
    return dict(sample_sec_per_frame=sec, fps_sample=1.0 / sec, fps_scaled=ratio / sec, cores=torch.get_num_threads(),
                sample="oracle port (== reference PyTorch fp32 path) on one %dx%d pair, %d warm-up + %d timed; frames/s scaled by the pixel ratio "
                       "%d/%d to the 1088x1920 workload (all-pairs corr term grows faster, so this flatters the CPU)"
                       % (SAMPLE_H, SAMPLE_W, warmup, steps, SAMPLE_H * SAMPLE_W, H_PAD * W_PAD))


def run_reference(args, rank):
    if rank != 0:
        return
    r = cpu_reference_fps(max(1, args.steps), max(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["fps_scaled"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * r["sample_sec_per_frame"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1 x 1920x1080 pair (padded 1088x1920), t=0.5, GIMM-VFI-R, random-init weights", "timed_sample": "%dx%d" % (SAMPLE_H, SAMPLE_W)},
        "cpu_baseline": {"value": r["fps_scaled"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["fps_scaled"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--height", type=int, default=H_PAD)
    ap.add_argument("--width", type=int, default=W_PAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="mixed", choices=["mixed", "mixed4", "3xtf32", "tf32", "fp32"],
                    help="mixed (default): RAFT + correlation on tcgen05 with 3xTF32 operand splitting and register-promoted "
                         "accumulation, post-RAFT convs on tcgen05 TF32, the final decoder's 256-channel residual trunk stored in fp16 on "
                         "tcgen05 kind::f16; 3xtf32: the trunk in TF32 too; tf32: RAFT on fp32 CUDA cores instead; fp32: everything on "
                         "fp32 CUDA cores.  All meet max|d imgt_pred| <= 1e-3 vs the reference (profiles/).  mixed4: experimental mode 4, see DESIGN.md.")
    ap.add_argument("--profile-json", default="", help="write the per-kernel CUDA-event breakdown here")
    ap.add_argument("--timesteps", type=int, default=1,
                    help="T interpolated frames per pair at t = i/(T+1): 1 = the headline metric (t=0.5); 7 = the reference's N=8 video setting "
                         "(flow estimation amortised over 7 frames; secondary figure, profiles/)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from gimmvfi_b200.parallel import init_from_env

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return

    import torch.distributed as dist

    rank, local, world = init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from gimmvfi_b200 import GIMMVFI_R
    from gimmvfi_b200.synth import synth_pair

    H, W, B, T, tval = args.height, args.width, 1, max(1, args.timesteps), 0.5
    tvals = [0.5] if T == 1 else [i / (T + 1) for i in range(1, T + 1)]
    model = GIMMVFI_R(seed=0).to(dev).eval()
    model.tensor_cores = {"fp32": 0, "tf32": 1, "3xtf32": 2, "mixed": 3, "mixed4": 4}[args.precision]
    xs_host = synth_pair(H, W, seed=100 + rank).pin_memory()
    xs = xs_host.to(dev, non_blocking=True)
    coord = [(model.sample_coord_input(B, (H, W), [tv], device=dev), None) for tv in tvals]
    tt = [tv * torch.ones(B, device=dev) for tv in tvals]
    gathered = torch.empty(world * B * T, 3, H, W, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        out = model(xs, coord, t=tt)
        img = torch.stack(out["imgt_pred"], 0).reshape(-1, 3, H, W) if T > 1 else out["imgt_pred"][0]
        if world > 1:
            dist.all_gather_into_tensor(gathered, img.contiguous())   # the single output collective
        return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- device-timed steps (CUDA events on the launching stream; L2 flushed between steps)
    evs = []
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    launches = model.engine.last_launches
    # ---- end-to-end through the public API: pinned host input -> H2D -> forward -> D2H of the frame, every step.
    # The loop is what a video caller runs (src/video_Nx.py:134-216 of the reference walks consecutive pairs): the copies of
    # step i+1 / i-1 travel on a second stream while step i computes; nothing is reused across steps and the whole K-step
    # region (all copies included) is timed by the wall clock between two full synchronisations.
    out_host = [torch.empty(B * T, 3, H, W).pin_memory() for _ in range(2)]
    x_dev = [torch.empty_like(xs) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]     # H2D of the buffer landed
    ev_free = [torch.cuda.Event() for _ in range(2)]   # forward that read the buffer finished
    ev_out = [torch.cuda.Event() for _ in range(2)]    # D2H of the result buffer finished
    keep = [None, None]

    def h2d(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[i % 2])
            x_dev[i % 2].copy_(xs_host, non_blocking=True)
            ev_in[i % 2].record(copy_stream)

    barrier()
    for e in ev_free + ev_out:
        e.record(main_stream)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    h2d(0)
    for i in range(args.steps):
        main_stream.wait_event(ev_in[i % 2])
        if i + 1 < args.steps:
            h2d(i + 1)
        c = [(model.sample_coord_input(B, (H, W), [tv], device=dev), None) for tv in tvals]
        o = model(x_dev[i % 2], c, t=[tv * torch.ones(B, device=dev) for tv in tvals])
        img = torch.stack(o["imgt_pred"], 0).reshape(-1, 3, H, W) if T > 1 else o["imgt_pred"][0]
        if world > 1:
            dist.all_gather_into_tensor(gathered, img.contiguous())
        ev_free[i % 2].record(main_stream)
        keep[i % 2] = img
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[i % 2])
            copy_stream.wait_event(ev_out[i % 2])
            img.record_stream(copy_stream)
            out_host[i % 2].copy_(img, non_blocking=True)
            ev_out[i % 2].record(copy_stream)
    torch.cuda.synchronize(dev)
    e2e_t = [(time.perf_counter() - t0) / args.steps]
    barrier()
    clocks = sampler.stop() if sampler else None
    e2e_ms = 1000.0 * sum(e2e_t) / len(e2e_t)
    # ---- per-kernel breakdown (CUDA events around every launch of one extra step)
    model.engine.set_profile(True)
    step()
    prof = model.engine.profile()
    model.engine.set_profile(False)
    # max over ranks
    if world > 1:
        tm = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms, e2e_ms = tm.tolist()
    if rank == 0:
        peaks = load_peaks()
        raw_prof = prof
        agg = {}
        for k, v in prof.items():   # labels carry the layer shape; aggregate per kernel for the headline
            a = agg.setdefault(k.split(" ")[0], {"ms": 0.0, "work": 0.0, "launches": 0})
            a["ms"] += v["ms"]; a["work"] += v["work"]; a["launches"] += v["launches"]
        prof = agg
        total_ms = sum(v["ms"] for v in prof.values()) or 1.0
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
        conv_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("conv2d"))
        conv_fl = sum(v["work"] for k, v in prof.items() if k.startswith("conv2d"))
        dname, d = dom
        is_flop = dname.startswith("conv2d") or dname.startswith("corr_gemm")
        # ncu --set full captures of the dominant LAYER SHAPE of each tensor-core kernel (profiles/r01_ncu_*.jsonl):
        # dram__bytes_read.sum + dram__bytes_write.sum per launch
        NCU_TRAFFIC = {"conv2d_tc_f16": {"layer": "3x3 256->256 @1088x1920, fp16 storage (profiles/r01_ncu_conv2d_tc_f16_256x256_1080p.jsonl)",
                                         "bytes": 1.072938e9 + 1.031320e9, "algorithmic_bytes": 2 * 1088 * 1920 * 256 * 2.0},
                       "conv2d_tc_tf32": {"layer": "3x3 256->256 @1088x1920", "bytes": 2.146013e9 + 2.092577e9, "algorithmic_bytes": 2 * 1088 * 1920 * 256 * 4.0},
                       "conv2d_tc_3xtf32": {"layer": "1x5 384->128 @2x136x240 (SepConvGRU gate; profiles/r01_ncu_conv2d_tc_3xtf32_gru_384x128_pair.jsonl)", "bytes": 102.902784e6 + 12.797184e6,
                                            "algorithmic_bytes": 2 * 136 * 240 * (384 + 128) * 4.0}}
        NOTES = {"conv2d_tc_f16": "tcgen05 kind::f16 implicit GEMM on the fp16-stored residual trunk (TMA halo tiles, TMEM fp32 accumulators)",
                 "conv2d_tc_tf32": "tcgen05 kind::tf32 implicit GEMM (TMA halo tiles, TMEM accumulators); TF32 peak is half the bf16 peak used as denominator",
                 "conv2d_tc_3xtf32": "tcgen05 3xTF32 (3 MMAs per K step + register-promoted accumulation): 'achieved' counts ALGORITHMIC flops, "
                                     "the tensor pipe executes 3x that at the TF32 rate (= bf16 peak / 2), so the ceiling of this ratio is 1/6; small-resolution RAFT layers on CTA pairs with TMEM-resident split operands (ncu: tensor pipe 58.8 % active)",
                 "conv2d_simt_n64": "fp32 CUDA-core implicit GEMM measured against the tensor-pipe peak (the layer class is tensor-bound, SURVEY 8(d))"}
        if is_flop:
            ach = d["work"] / (d["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": dname, "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sust"],
                    "traffic": NCU_TRAFFIC.get(dname, {}).get("bytes"), "traffic_detail": NCU_TRAFFIC.get(dname),
                    "peak_source": "%s bf16 sustained (MEASURED_PEAKS.json)" % peaks["which"],
                    "launches": d["launches"], "avg_launch_ms": d["ms"] / d["launches"], "share_of_step": d["ms"] / total_ms,
                    "note": NOTES.get(dname, "")}
            # the single most expensive layer shape, for which the ncu capture above was taken
            top = max(((k, v) for k, v in raw_prof.items() if k.startswith(("conv2d", "corr_gemm"))), key=lambda kv: kv[1]["ms"])
            roof["top_layer"] = {"name": top[0], "ms_total": top[1]["ms"], "launches": top[1]["launches"],
                                 "tflops": top[1]["work"] / (top[1]["ms"] * 1e-3) / 1e12, "frac_of_peak": top[1]["work"] / (top[1]["ms"] * 1e-3) / 1e12 / peaks["tf_sust"]}
        else:
            ach = 4.0 * d["work"] / (d["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": dname, "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None,
                    "launches": d["launches"], "share_of_step": d["ms"] / total_ms}
        if args.profile_json:
            with open(args.profile_json, "w") as f:
                json.dump({"per_kernel": prof, "per_layer": raw_prof, "sum_ms": total_ms, "step_ms": ms}, f, indent=1)
        P = H * W
        fl = flops_per_frame(P, T)
        line = {
            "metric": METRIC, "value": world * B * T / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "3xtf32": "tf32", "mixed": "tf32+f16 operands, f32 accumulate", "mixed4": "tf32+f16 operands, f32 accumulate"}[args.precision], "data": "synthetic",
            "config": {"precision": {"tf32": "RAFT fp32 (CUDA cores); post-RAFT convs TF32 tcgen05, fp32 accumulate; max|d imgt_pred| vs CPU reference 5.4e-4 at this size (profiles/r01_parity_1080p.log)", "fp32": "fp32 everywhere", "mixed": "RAFT + correlation: tcgen05 3xTF32 (operand split, fp32 register-promoted accumulation); post-RAFT convs: tcgen05 TF32; final-decoder residual trunk (256 ch): fp16 storage + tcgen05 kind::f16; fp32 accumulate everywhere; max|d imgt_pred| vs CPU reference 5.9e-4 at this size, PSNR 81.2 dB (profiles/r01_parity_1080p_mode3_final.log)", "mixed4": "EXPERIMENTAL precision mode 4 (mixed + fp16 storage of the 32/64-channel full-resolution chains, HypoNet activations, init-decoder trunk and the decoder concat); parity validated on the CPU emulation only (DESIGN.md)", "3xtf32": "RAFT + correlation: tcgen05 3xTF32 (operand split, fp32 register-promoted accumulation); post-RAFT convs: tcgen05 TF32, fp32 accumulate; max|d imgt_pred| vs CPU reference 5.5e-4 at this size, PSNR 81.7 dB (profiles/r01_parity_1080p_modes.log)"}[args.precision], "workload": "%d x 1920x1080 pair per GPU (padded %dx%d), %s, GIMM-VFI-R (RAFT 20 iters), random-init weights, all reference outputs produced"
                                   % (B, H, W, "t=0.5, T=1" if T == 1 else "T=%d frames per pair at t=i/%d" % (T, T + 1)), "parallelism": "pairs sharded, 1 all-gather of output frames" if world > 1 else "single GPU",
                       "l2": "256 MiB L2 flush between timed steps; per-step working set ~30 GB >> L2",
                       "algorithmic_tflop_per_frame": fl / 1e12, "achieved_tflops_end_to_end": world * fl / (ms * 1e-3) / 1e12},
            "e2e": {"value": world * B * T / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": xs_host.numel() * 4,
                    "d2h_bytes_per_step": out_host[0].numel() * 4,
                    "how": "K-step loop, wall clock between full syncs; per-step H2D/D2H on a copy stream overlapped with the previous/next forward"},
            "gpu_launches": int(launches) * args.steps,
            "launches_per_step": int(launches),
            "roofline": roof,
            "kernel_shares": {k: round(v["ms"] / total_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]},
            "conv_tflops": conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms else None,
            "clocks": clocks,
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded: the CPU arm runs in a child process with a hard time limit
            try:
                o = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                                   capture_output=True, text=True, timeout=240).stdout.strip().splitlines()
                cb = json.loads([l for l in o if l.startswith("{")][-1])["cpu_baseline"]
                cb["sample_sec_per_frame"] = 1.0 / (cb["value"] * (H_PAD * W_PAD) / float(SAMPLE_H * SAMPLE_W))
                line["cpu_baseline"] = cb
            except Exception as ex:  # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cpu_threads(), "kind": "port", "sample": "CPU arm did not finish within 240 s: %r" % (ex,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
