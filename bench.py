#!/usr/bin/env python
"""bench.py -- aggregate input MS/s of the SDRReceiver channelizer hot path (25E plan).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = `--blocks` callbacks (default 4 = one second of signal) for each of `--streams`
independent 1.536 MS/s uint8-IQ streams per GPU (default 128: the 1024-stream configuration
of BASELINE.json sharded over 8 GPUs; weak scaling, no data-path collective). Rank 0 prints
ONE JSON line:
  value     whole-job input MS/s with the uint8 IQ already resident in HBM (kernels only)
  e2e       same metric through the host-facing C-ABI call sdrb_bank_process_host: pinned
            host uint8 in, H2D, kernels, D2H, pinned int16 out, every step
  roofline  the kernel class with the largest share of the step, algorithmic bytes / its
            CUDA-event time vs the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the unmodified reference (oracle/_ref) on this box's host cores, bounded sample
`--impl reference` times the reference's own CPU code instead (same metric/config).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aggregate input MS/s (uint8 IQ, 25E plan)"
HBM_FALLBACK_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--plan", default="25E")
    ap.add_argument("--streams", type=int, default=128, help="streams per GPU")
    ap.add_argument("--blocks", type=int, default=4, help="callbacks per stream per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------
# synthetic input: a few fully synthesised base streams, the rest are time-rotated copies
# --------------------------------------------------------------------------------------
def base_streams(plan, n_base, n_samples):
    from sdrreceiver_b200 import synth
    car = synth.carriers_for_plan(plan.center, plan.subs)
    return [synth.make_iq(plan.fs, n_samples, car, stream=s) for s in range(n_base)]


def bind_to_gpu_cpus(index):
    """Pin this process to the CPU cores NVML reports as local to GPU `index`, so that the pinned staging
    buffers of the end-to-end leg live on that GPU's NUMA node (matters from 2 ranks up). Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        getr = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = getr(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)      # NVML queries take driver locks: keep them rare next to ~130 launches per step

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference, one process per host core
# --------------------------------------------------------------------------------------
def run_reference_cpu(plan, plan_path, cores, blocks_timed, skip):
    """Runs `cores` copies of oracle/_ref/sdr_ref_i16 (one stream each) and returns
    (total_samples, seconds = slowest process's time inside demodData)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "sdr_ref_i16")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/sdr_ref_i16 missing (built from /root/reference by oracle/Makefile)")
    n_blocks = blocks_timed + skip
    base = base_streams(plan, 1, plan.block * n_blocks)[0]
    with tempfile.TemporaryDirectory() as d:
        procs = []
        for c in range(cores):
            fn = os.path.join(d, "s%d.u8" % c)
            np.roll(base, 2 * 977 * c).tofile(fn)
        for c in range(cores):
            procs.append(subprocess.Popen([exe, "--ini", plan_path, "--in", os.path.join(d, "s%d.u8" % c), "--time",
                                           "--skip", str(skip), "--blocks", str(n_blocks)],
                                          stdout=subprocess.PIPE, text=True))
        total, worst = 0, 0.0
        for p in procs:
            out = p.communicate()[0].split()
            if p.returncode != 0 or len(out) != 2:
                raise RuntimeError("reference harness failed")
            total += int(out[0])
            worst = max(worst, float(out[1]))
    return total, worst


def reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    plan_path = os.path.join(ROOT, "plans", args.plan + ".ini")
    from sdrreceiver_b200 import binding as B
    plan = B.Plan(plan_path)                 # plan arithmetic only; no GPU is touched on this arm
    op = {"Fs": plan.fs, "block": plan.block}
    cores = os.cpu_count() or 1
    # each step = every core pushes `--blocks` callbacks of its own stream through the reference
    total, secs = run_reference_cpu(plan, plan_path, cores, args.blocks * args.steps, args.blocks * args.warmup)
    value = total / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "realtime_x": value * 1e6 / op["Fs"],
        "config": {"workload": "%s plan, %d host processes x 1 stream, %d callbacks per step" % (args.plan, cores, args.blocks),
                   "plan": args.plan, "streams": cores, "blocks_per_step": args.blocks},
        "cpu_baseline": {"value": value, "unit": "MS/s", "cores": cores, "kind": "reference",
                         "sample": "%d processes x %d timed callbacks (g++ -O2 -ffp-contract=off build of the unmodified reference)" % (cores, args.blocks * args.steps)},
        "e2e": {"value": value, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def b200_arm(args):
    import torch
    import torch.distributed as dist
    from sdrreceiver_b200 import binding as B, shard

    rank, local_rank, world = dist_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    numa = bind_to_gpu_cpus(local_rank)          # before the pinned buffers are allocated (first touch)
    plan_path = os.path.join(ROOT, "plans", args.plan + ".ini")
    plan = B.Plan(plan_path)
    S, NB, Bk = args.streams, args.blocks, plan.block
    row = NB * Bk * 2
    samples_per_step = S * NB * Bk

    # ---- input: stream s of rank r is global stream r + world*s (independent units, no exchange)
    n_base = 2
    base = base_streams(plan, n_base, NB * Bk)
    pin_in = B.PinnedBuffer(S * row)
    h_iq = pin_in.array.reshape(S, row)
    for s, g in enumerate(shard.stream_ids(rank, world, S)):
        h_iq[s] = np.roll(base[g % n_base], 2 * 977 * g)
    d_iq = torch.from_numpy(h_iq).to(dev)                      # resident copy for the kernel metric
    d_pcm = torch.empty((S, NB, plan.pcm_per_block), dtype=torch.int16, device=dev)
    pin_out = B.PinnedBuffer(S * NB * plan.pcm_per_block * 2)
    # the end-to-end leg keeps two calls in flight: a second pair of pinned host buffers
    pin_in2 = B.PinnedBuffer(S * row)
    pin_in2.array[:] = pin_in.array
    pin_out2 = B.PinnedBuffer(S * NB * plan.pcm_per_block * 2)

    bank = B.Bank(plan, S, NB, device=local_rank)
    stream = torch.cuda.Stream(device=dev)

    # the input is resident before the timed region: tell the library so with an event (a streaming
    # caller would record it after the copy that fills its next input buffer)
    torch.cuda.synchronize()
    in_ready = torch.cuda.Event()
    in_ready.record(stream)
    in_ready.synchronize()

    def step_device():
        bank.process_device(d_iq.data_ptr(), row, NB, d_pcm.data_ptr(), None, stream.cuda_stream, in_ready.cuda_event)

    host_bufs = [(pin_in.ptr, pin_out.ptr), (pin_in2.ptr, pin_out2.ptr)]
    host_step = [0]

    def step_host():
        # sdrb_bank_process_host_async: every step copies its own input from pinned host memory and its own result
        # back; at most two steps are in flight, so step i+1's copy-in overlaps step i's filters and copy-out
        i, o = host_bufs[host_step[0] & 1]
        host_step[0] += 1
        bank.process_host_async(i, row, NB, o, None)
        bank.host_wait(1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- kernel metric: inputs resident in HBM (393 MB per step > 126 MB L2) ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    launches = bank.last_launches * args.steps
    dev_ms = e0.elapsed_time(e1)
    # second pass over the same steps with a CUDA-event pair around every launch (on the stream it
    # is launched on): per-kernel durations for the roofline, kept out of the number above
    bank.set_timing(True)
    for _ in range(args.steps):
        step_device()
    torch.cuda.synchronize()
    ktimes = bank.kernel_times()
    bank.set_timing(False)

    # ---- end to end through the host-facing call ----
    e2e_ms = None
    if not args.no_e2e:
        for _ in range(args.warmup):
            step_host()
        bank.host_wait(0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        bank.host_wait(0)                                      # the last result is on the host
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # timings + output digest of every rank, gathered over NCCL (the only collective in the job)
    digest = shard.pcm_digest(d_pcm.cpu().numpy())
    stats_all, digests = shard.gather([dev_ms, e2e_ms if e2e_ms is not None else 0.0, float(samples_per_step)],
                                      digest, device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    agg = shard.aggregate(stats_all, args.steps)
    max_dev_ms, max_e2e_ms, total_samples = agg["dev_ms"], agg["e2e_ms"], agg["total_samples"]
    value = agg["value_msps"]
    e2e_value = agg["e2e_msps"] if e2e_ms is not None else None

    # ---- roofline of the dominant kernel class (rank 0's events) ----
    fs = plan.fs
    main_out_b = 8.0 * sum(m["out_rate"] for m in plan.mains) / fs
    z_b = 8.0 * sum(s["out_rate"] * (s["late"] or 1) for s in plan.subs) / fs
    d_b = 8.0 * sum(s["out_rate"] for s in plan.subs if s["late"]) / fs
    usb_in_b = 8.0 * sum(s["out_rate"] for s in plan.subs) / fs
    pcm_b = 2.0 * sum(s["out_rate"] for s in plan.subs) / fs
    alg_bytes = {                                   # algorithmic bytes per input complex sample
        "dc_scan": 2.0,
        "ingest_main": 2.0 + main_out_b,
        "sub_cascade": main_out_b + z_b,
        "late_fir": z_b if d_b else 0.0,
        "usb_audio": usb_in_b + pcm_b,
        "carry": 0.0,
    }
    peak, peak_src = hbm_peak()
    per_step_ms = {k: v[0] / args.steps for k, v in ktimes.items()}      # sum of that class's launches in one step
    launches_per_step = {k: v[1] / args.steps for k, v in ktimes.items()}
    dom = max((k for k in per_step_ms if k != "dc_scan"), key=lambda k: per_step_ms[k])   # dc_scan runs on the side stream
    dom_ms = per_step_ms[dom]
    achieved = alg_bytes[dom] * samples_per_step / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    clocks = sampler.summary()
    sm_mhz = clocks["sm_mhz"] or 1965.0
    fp32_nominal = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    # FP32 denominator: an FMA loop measured on this device after the timed region (scalar FFMA and packed
    # FFMA2 chains, the better of the two); the nominal figure is kept beside it
    fp32_probe = {"ffma": B.probe_fp32_tflops(False), "ffma2": B.probe_fp32_tflops(True)}
    fp32_peak = max(fp32_probe.values())
    fp32_src = ("measured FMA loop (sdrb_probe_fp32_tflops): FFMA %.1f, FFMA2 %.1f TFLOP/s; nominal 148 SM x 128 lanes x 2 x "
                "%.0f MHz = %.1f" % (fp32_probe["ffma"], fp32_probe["ffma2"], sm_mhz, fp32_nominal))
    step_ms = max_dev_ms / args.steps
    # DRAM traffic of that kernel class from the committed ncu --set full capture (same plan and bank
    # size), per "launch" = the class's launches of one callback, like `achieved`
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tj = json.load(f)
        if tj["config"]["plan"] == args.plan and tj["config"]["streams_per_gpu"] == S:
            traffic, traffic_src = tj["per_class_per_callback"].get(dom), "profiles/r01_traffic.json (ncu --set full)"
    except Exception:
        pass
    samples_per_callback = S * Bk
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "launch": "the %s kernels of one callback (%d streams x %d samples)" % (dom, S, Bk),
        "alg_bytes_per_launch": alg_bytes[dom] * samples_per_callback, "launch_ms": dom_ms / NB,
        "kernel_ms_per_step": dom_ms, "share_of_step": dom_ms / step_ms if step_ms > 0 else None,
        "alg_bytes_per_sample": alg_bytes[dom],
    }
    # the same kernel against the FP32 roofline (it is FP32-issue bound): algorithmic
    # flops by SURVEY.md 8(d)'s counting rule
    hb = lambda decim: sum(20.0 / 2 ** a for a in range(1, decim + 1))
    alg_flops = {
        "ingest_main": 10.0 + sum(6.0 + hb(m["decim"]) for m in plan.mains),
        "sub_cascade": sum((s["Fs"] / fs) * (6.0 + hb(s["decim"])) for s in plan.subs),
        "usb_audio": sum((s["out_rate"] / fs) * (2.0 * 62 + 1 + 2.0 * s["n_lpf_taps"] + 2.0) for s in plan.subs),
    }
    if dom in alg_flops and dom_ms > 0:
        tf = alg_flops[dom] * samples_per_step / (dom_ms * 1e-3) / 1e12
        roofline["fp32"] = {"achieved_tflops": tf, "peak_tflops": fp32_peak, "frac": tf / fp32_peak,
                            "alg_flops_per_sample": alg_flops[dom], "peak_source": fp32_src}
    line = {
        "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s plan x %d streams per GPU (%d total) x %d callbacks (%.2f s of signal) per step; "
                               "input %.0f MB per GPU per step, larger than L2" % (
                                   args.plan, S, S * world, NB, NB / plan.bufsplit, S * row / 1e6),
                   "plan": args.plan, "streams_per_gpu": S, "blocks_per_step": NB, "sample_rate": fs,
                   "l2_policy": "input larger than L2 (no flush)", "parallelism": "streams sharded, no collective",
                   "cpus_bound_per_rank": numa},
        "realtime_x": value * 1e6 / fs,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "MS/s", "h2d_bytes_per_step": S * row,
                "d2h_bytes_per_step": S * NB * plan.pcm_per_block * 2, "ms_per_step": max_e2e_ms / args.steps,
                "timing": "host wall clock around K sdrb_bank_process_host_async calls, two in flight, last result waited for "
                          "(pinned host buffers in and out every step)",
                "realtime_x": (e2e_value * 1e6 / fs) if e2e_value else None} if e2e_ms is not None else None,
        "gpu_launches": launches,
        "roofline": roofline,
        "roofline_pipeline": {"bound": "hbm", "achieved": plan.alg_bytes * samples_per_step / (step_ms * 1e-3) / 1e9,
                              "peak": peak, "unit": "GB/s",
                              "frac": plan.alg_bytes * samples_per_step / (step_ms * 1e-3) / 1e9 / peak,
                              "alg_bytes_per_sample": plan.alg_bytes,
                              "fp32": {"achieved_tflops": plan.alg_flops * samples_per_step / (step_ms * 1e-3) / 1e12,
                                       "peak_tflops": fp32_peak,
                                       "frac": plan.alg_flops * samples_per_step / (step_ms * 1e-3) / 1e12 / fp32_peak,
                                       "alg_flops_per_sample": plan.alg_flops,
                                       "peak_source": fp32_src}},
        "kernels_ms_per_step": per_step_ms,
        "kernel_launches_per_step": launches_per_step,
        "digests": digests,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            total, secs = run_reference_cpu(plan, plan_path, cores, 12, 1)
            line["cpu_baseline"] = {"value": total / secs / 1e6, "unit": "MS/s", "cores": cores, "kind": "reference",
                                    "sample": "%d processes x 12 timed callbacks (3 s of signal each), unmodified reference "
                                              "built g++ -O2 -ffp-contract=off" % cores}
        except Exception as ex:   # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "MS/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % ex}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
