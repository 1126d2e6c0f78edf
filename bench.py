#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (BASELINE.json metric).

One step = one pass of the hot path over one synthetic ComplexF32 stream shard:
    y = conv(x, v)              overlap-save FFT convolution, 4097-tap FIR on 2^26 samples   (BASELINE config 2)
    P = welch_pgram(y[:2^26])   n = nfft = 4096, 50 % overlap, hanning, two-sided            (Welch stage of the metric)
metric = input samples per second through both stages (Gsamples/s), whole job over all ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 26]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N

Multi-GPU (weak scaling): every rank owns a 2^26-sample shard of one long stream (plus the nv-1 sample left halo),
convolves its own output range with no collective, accumulates the Welch power of its own segments scaled by the
GLOBAL 1/(k r), and the only exchange is one NCCL all-reduce of the 4096-bin power vector per step.

`--impl reference` times the reference's CPU path.  Julia/FFTW are not available in this image, so it runs the CPU
oracle port (oracle/, scipy pocketfft with all host threads) of the same two stages on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NV = 4097
NSEG, NOVERLAP = 4096, 2048
METRIC = "Gsamples/s filt+welch on 2^26 cplx-F32"


def make_taps():
    """4097-tap complex bandpass: Hamming-windowed sinc shifted to 0.3 pi."""
    n = np.arange(NV) - NV // 2
    return (0.2 * np.sinc(0.2 * n) * np.hamming(NV) * np.exp(1j * np.pi * 0.3 * n)).astype(np.complex64)


def hanning64(n):
    x = -0.5 + np.arange(n, dtype=np.float64) / (n - 1)
    return 0.5 * (1 + np.cos(2 * np.pi * x))


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------- CPU baseline (oracle port)

def cpu_reference_step(x, v, win, workers):
    """The reference's CPU path for one step, restated (oracle/): overlap-save conv with the reference's own block
    length (optimalfftfiltlength -> 65536, src/dspbase.jl:268-291, 490-609) + welch_pgram (src/periodograms.jl:746-759),
    Float32 arithmetic, pocketfft with `workers` threads."""
    import scipy.fft as sfft
    from oracle import dspbase as od
    from oracle import periodograms as op
    with sfft.set_workers(workers):
        nfft = od.optimalfftfiltlength(len(v), len(x))
        y = od.conv_kern_os(x, v, nfft, batched=True)
        p, _ = op.welch_pgram(y[:len(x)], NSEG, NOVERLAP, onesided=False, nfft=NSEG, window=win)
    return y, p


def time_cpu_baseline(log2_sample, reps, workers):
    rng = np.random.default_rng(1002)
    n = 1 << log2_sample
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2)).astype(np.complex64)
    v = make_taps()
    win = hanning64(NSEG)
    cpu_reference_step(x[: 1 << 18], v, win, workers)      # warm-up (plan caches, imports)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_step(x, v, win, workers)
        ts.append(time.perf_counter() - t0)
    return n, ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    log2s = min(args.log2n, 23)
    # warm-up + K timed steps, each a bounded sample of 2^23 samples (~1-3 s of CPU work per step)
    n, ts = time_cpu_baseline(log2s, max(1, args.steps), workers)
    ms = 1e3 * float(np.mean(ts))
    val = n / (ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": val, "unit": "Gsamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64 (ComplexF32)",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": f"overlap-save conv 4097-tap + welch_pgram(4096, 50%, hanning) on 2^{log2s} ComplexF32 "
                               "(bounded sample of the 2^26 workload)", "nfft_conv": 65536},
        "cpu_baseline": {"value": val, "unit": "Gsamples/s", "cores": workers, "kind": "port",
                         "sample": f"2^{log2s} samples per step; oracle port (numpy + scipy pocketfft, Float32), "
                                   "Julia/FFTW not installable in this image"},
        "e2e": {"value": val, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import torch
    import dspb200
    from dspb200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib.dspb200_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n = 1 << args.log2n                    # samples per rank
    n_global = n * world
    halo = NV - 1
    hop = NSEG - NOVERLAP

    # ---- synthetic input, resident in HBM before the timed region.  Rank r holds samples [r*n - halo, (r+1)*n + tail)
    # of the global stream (tail = the Welch overlap its last segments need from the right neighbour's range).
    g = torch.Generator(device=dev)
    k_global = (n_global - NSEG) // hop + 1
    seg_begin = (rank * n + hop - 1) // hop if rank > 0 else 0          # segments whose start lies in this rank's range
    seg_end = min(k_global, ((rank + 1) * n + hop - 1) // hop)
    need_hi = (seg_end - 1) * hop + NSEG if seg_end > seg_begin else (rank + 1) * n
    hi = max((rank + 1) * n, min(need_hi, n_global))
    lo = max(0, rank * n - halo)
    # deterministic per-block generation so overlapping halos agree across ranks
    blk = 1 << 20
    x = torch.empty(hi - lo, dtype=torch.complex64, device=dev)
    for b0 in range((lo // blk) * blk, hi, blk):
        g.manual_seed(1002 + b0 // blk)
        chunk = torch.view_as_complex(torch.randn(blk, 2, generator=g, device=dev, dtype=torch.float32)) * (2 ** -0.5)
        s0, s1 = max(b0, lo), min(b0 + blk, hi)
        x[s0 - lo: s1 - lo] = chunk[s0 - b0: s1 - b0]
    taps = make_taps()
    win = hanning64(NSEG)
    norm2 = float(np.sum(win * win))
    r = k_global * 1.0 * norm2                                            # r = k * fs * norm2 (src/periodograms.jl:751)

    os_plan = _lib.OsPlan(taps, args.nfft)
    spec = _lib.SpecPlan(np.complex64, NSEG, NOVERLAP, NSEG, False, win)
    # the conv of the global stream restricted to this rank's own sample range (same-length filter output)
    out_lo, out_cnt = rank * n, hi - rank * n
    y = torch.empty(out_cnt, dtype=torch.complex64, device=dev)
    pw = torch.zeros(NSEG, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step():
        os_plan.exec_range_dev(x.data_ptr(), lo, x.numel(), y.data_ptr(), out_lo, out_cnt, sp)
        spec.welch_range_dev(y.data_ptr(), out_cnt, out_lo, seg_begin, seg_end, r, pw.data_ptr(), sp)
        if world > 1:
            dist.all_reduce(pw)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up through the device-timed and end-to-end regions
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()

    # ---- timed region: K steps, CUDA events on the launching stream; per-stage events for the roofline
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    l0 = _lib.launch_count()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e_start.record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        os_plan.exec_range_dev(x.data_ptr(), lo, x.numel(), y.data_ptr(), out_lo, out_cnt, sp)
        ev[i][1].record(stream)
        spec.welch_range_dev(y.data_ptr(), out_cnt, out_lo, seg_begin, seg_end, r, pw.data_ptr(), sp)
        if world > 1:
            dist.all_reduce(pw)
        ev[i][2].record(stream)
    e_stop.record(stream)
    sync_all()
    launches = _lib.launch_count() - l0
    total_ms = e_start.elapsed_time(e_stop)
    conv_ms = float(np.mean([ev[i][0].elapsed_time(ev[i][1]) for i in range(args.steps)]))
    welch_ms = float(np.mean([ev[i][1].elapsed_time(ev[i][2]) for i in range(args.steps)]))
    if world > 1:
        t = torch.tensor([total_ms, conv_ms, welch_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, conv_ms, welch_ms = [float(v) for v in t.tolist()]
    ms_per_step = total_ms / args.steps
    value = n_global / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the repo's public API (dspb200.conv / dspb200.welch_pgram, the mirror of the reference's
    # calls): every step copies the step's input from PINNED host memory to the GPU, filters, estimates the PSD and
    # reads the PSD back.  Pipeline form: the filter output stays in HBM between the two calls (DeviceArray), so the
    # stream crosses PCIe once.  `e2e_host_calls` is the same step through the two host-pointer C-ABI calls
    # (dspb200_os_exec + dspb200_welch_exec), where the filter output comes back to the host and is uploaded again.
    e2e = None
    e2e_host = None
    if not args.no_e2e:
        xh = torch.empty(n, dtype=torch.complex64).pin_memory()
        xh.copy_(x[(rank * n - lo): (rank * n - lo) + n].cpu())
        yh = torch.empty(n, dtype=torch.complex64).pin_memory()
        ph = torch.empty(NSEG, dtype=torch.float32).pin_memory()
        k_local = (n - NSEG) // hop + 1
        r_local = k_local * norm2
        reps = max(2, min(args.steps, 5))

        def timed(fn):
            fn()
            sync_all()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            dt_ = (time.perf_counter() - t0) / reps
            if world > 1:
                t_ = torch.tensor([dt_], device=dev, dtype=torch.float64)
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                dt_ = float(t_.item())
            return dt_

        wcfg = dspb200.WelchConfig(n, np.complex64, n=NSEG, noverlap=NOVERLAP, onesided=False, nfft=NSEG, window=win)
        xd = dspb200.DeviceArray((n,), np.complex64)
        result = {}

        def e2e_step():
            xd.copy_from_host_ptr(xh.data_ptr(), n * 8)                                  # H2D, pinned
            yd = dspb200.conv(xd, taps, algorithm="fft_overlapsave", nfft=(args.nfft or None))
            result["p"] = dspb200.welch_pgram(yd[:n], wcfg).power                        # D2H of the PSD

        dt = timed(e2e_step)
        e2e = {"value": n_global / dt / 1e9, "unit": "Gsamples/s", "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(n * 8), "d2h_bytes_per_step": int(NSEG * 4),
               "note": "public API pipeline: to-device copy from pinned host memory -> dspb200.conv -> dspb200.welch_pgram -> PSD to host"}

        def e2e_host_step():
            os_plan.exec_ptr(xh.data_ptr(), n, 1, yh.data_ptr(), n)             # filt-style same-length output
            spec.welch_ptr(yh.data_ptr(), n, r_local, ph.data_ptr())

        dth = timed(e2e_host_step)
        e2e_host = {"value": n_global / dth / 1e9, "unit": "Gsamples/s", "ms_per_step": dth * 1e3,
                    "h2d_bytes_per_step": int(2 * n * 8), "d2h_bytes_per_step": int(n * 8 + NSEG * 4),
                    "note": "two host-pointer C-ABI calls (dspb200_os_exec + dspb200_welch_exec), pinned buffers, chunked copy/compute overlap"}
        del xd

    clk = clocks.stop() if rank == 0 else None

    # ---- Welch on a real Float32 stream (BASELINE config 3) -- reported beside the headline, rank 0, N = 1 only
    extra = {}
    if world == 1 and not args.no_extra:
        xr = torch.randn(n, device=dev, dtype=torch.float32)
        spec_r = _lib.SpecPlan(np.float32, NSEG, NOVERLAP, NSEG, True, win)
        pr = torch.zeros(NSEG // 2 + 1, dtype=torch.float32, device=dev)
        k3 = (n - NSEG) // hop + 1
        for _ in range(3):
            spec_r.welch_dev(xr.data_ptr(), n, k3 * norm2, pr.data_ptr(), sp)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            spec_r.welch_dev(xr.data_ptr(), n, k3 * norm2, pr.data_ptr(), sp)
        b.record(stream)
        torch.cuda.synchronize()
        ms3 = a.elapsed_time(b) / args.steps
        extra["welch_f32_config3"] = {"ms": ms3, "gsamples_s": n / (ms3 * 1e-3) / 1e9,
                                      "hbm_gbs_algorithmic": 4.0 * n / (ms3 * 1e-3) / 1e9}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    conv_bytes = 16.0 * out_cnt                       # 8 B read + 8 B written per ComplexF32 sample
    achieved = conv_bytes / (conv_ms * 1e-3) / 1e9
    welch_bytes = 8.0 * out_cnt
    cpu_workers = os.cpu_count() or 1
    cb = None
    if world == 1 and not args.no_cpu:
        ns, ts = time_cpu_baseline(min(args.log2n, 23), 2, cpu_workers)
        cb = {"value": ns / float(np.mean(ts)) / 1e9, "unit": "Gsamples/s", "cores": cpu_workers, "kind": "port",
              "sample": f"2^{min(args.log2n, 23)} samples x 2 reps of the same two stages; oracle port (numpy + scipy "
                        "pocketfft, Float32, nfft 65536 as the reference picks); Julia/FFTW not installable here"}
    line = {
        "metric": METRIC, "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c64 (ComplexF32; f32 arithmetic)", "data": "synthetic",
        "config": {"workload": f"conv overlap-save 4097-tap FIR + welch_pgram(n=nfft=4096, 50% overlap, hanning, two-sided) "
                               f"on 2^{args.log2n} ComplexF32 samples per GPU (BASELINE configs[1] + Welch stage)",
                   "samples_per_gpu": n, "nfft_conv": os_plan.nfft, "conv_fused": os_plan.fused,
                   "l2_policy": "inputs (512 MiB per stage) exceed the 126 MB L2; no explicit flush",
                   "parallelism": f"stream range-sharded over {world} GPU(s); NCCL all-reduce of the 4096-bin Welch power only"},
        "stages_ms": {"conv": conv_ms, "welch_plus_allreduce": welch_ms},
        "roofline": {"bound": "hbm", "kernel": "os_fused_kernel<float,16384,complex>" if os_plan.fused else "cuFFT pipeline",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": conv_bytes,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at 2^26 samples, from the committed
                     # `ncu --set full` capture (profiles/r1_conv_v3.txt); null for other sizes / the cuFFT path
                     "traffic": (537.194e6 + 491.504e6) if (args.log2n == 26 and os_plan.fused and world == 1) else None,
                     "traffic_source": "profiles/r1_conv_v6.txt (ncu --set full, one launch)",
                     "welch_stage": {"achieved": welch_bytes / (welch_ms * 1e-3) / 1e9, "frac": welch_bytes / (welch_ms * 1e-3) / 1e9 / peak,
                                     "algorithmic_bytes_per_launch": welch_bytes}},
        "cpu_baseline": cb, "e2e": e2e, "e2e_host_calls": e2e_host, "gpu_launches": int(launches), "clocks": clk, "extra": extra,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=26)
    ap.add_argument("--nfft", type=int, default=0, help="overlap-save block transform (0 = library choice)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
