#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: RX IQ Msamples/s, batched channels).

Workload (BASELINE.json configs[1]): 64 channels of 4FSK-2k RX (RRC + clock recovery + Viterbi), 1 Msps per
channel, per GPU (weak scaling: every rank runs its own 64 channels; channels are independent so there is
no data-path collective).  One step = one pass of the whole RX chain over [64][2^22] synthetic gr_complex
samples (4.19 s of air time per channel).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (device-resident `value`, host `e2e`)
  python bench.py --impl reference ...                     the CPU restatement of the reference chain (oracle port)

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHANNELS = 64
T_STEP = 1 << 22
D1 = 50
ALG_BYTES_PER_SAMPLE = 8.0 * (1.0 + 1.0 / D1)      # stage-1: read 8 B, write 8/D B per input sample
N_BASES = 8
WORKLOAD = "64ch 4FSK-2k-FM RX (make_gr_demod_4fsk(5,1e6,1700,3000,true)): /50 polyphase FIR + LPF + quad demod + RRC + symbol sync + CCSDS Viterbi + descrambler, 1 Msps/ch, T=2^22 samples/ch/step"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons = index, False, [], set()
        self.max_mhz = None
        self.armed = False                 # only samples taken while armed (the timed region) count
        self.ready = threading.Event()     # NVML is initialised: the first sample can be taken at once

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            self.ready.set()
            while not self.stop_flag:
                if not self.armed:
                    time.sleep(0.0005)
                    continue
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.001)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("sampler_error:%s" % type(e).__name__)
            self.ready.set()

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_chain_rate(seconds_target, n_threads, t_samples):
    """Times the CPU restatement of the reference chain (oracle port): n_threads channels in parallel,
    one channel per thread, each processing t_samples-sample chunks until ~seconds_target elapse."""
    from oracle import oracle as O
    from qradiolink_b200 import synth
    base = synth.burst_4fsk(1000, t_samples)
    rng = np.random.default_rng(1)
    xs = []
    for i in range(n_threads):
        noise = (rng.standard_normal(t_samples) + 1j * rng.standard_normal(t_samples)) * 0.04
        xs.append((0.8 * np.roll(base, 37 * i) + noise).astype(np.complex64))
    O.lib()
    done = [0] * n_threads
    t_end = [0.0]

    def worker(i):
        rx = O.Rx(O.DEMOD_4FSK, 5, 1000000, 1700, 3000, 1)
        while time.perf_counter() < t_end[0]:
            rx.work(xs[i])
            for p in (0, 1, 2):
                rx.port(p)
            done[i] += t_samples

    # warm-up pass
    t_end[0] = time.perf_counter() + 0.5
    th = [threading.Thread(target=worker, args=(i,)) for i in range(n_threads)]
    [t.start() for t in th]; [t.join() for t in th]
    done[:] = [0] * n_threads
    t0 = time.perf_counter()
    t_end[0] = t0 + seconds_target
    th = [threading.Thread(target=worker, args=(i,)) for i in range(n_threads)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    return sum(done) / dt / 1e6, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t_s = 1 << 20
    # each "step" = a bounded sample of the workload: all host threads demodulate for ~1.5 s
    rates = []
    for i in range(args.warmup + args.steps):
        r, dt = cpu_chain_rate(1.5, cores, t_s)
        if i >= args.warmup:
            rates.append((r, dt))
    v = float(np.mean([r for r, _ in rates]))
    line = {
        "impl": "reference", "metric": "RX IQ Msamples/s (batched channels)", "value": v, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean([dt for _, dt in rates]) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port of the reference GNU Radio chain (GNU Radio/VOLK not installable here); one channel per host thread"},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
                         "sample": "%d threads x 2^20-sample chunks of the 4FSK-2k-FM RX chain for ~1.5 s per step" % cores},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import qradiolink_b200 as q
    from qradiolink_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if q.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    C, T = CHANNELS, T_STEP
    bases = [synth.burst_4fsk(1000 + 97 * rank + i, T) for i in range(N_BASES)]
    X = synth.batch_on_device(bases, C, seed=4242 + rank, device=dev)
    torch.cuda.synchronize()

    blk = q.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T, device=local)
    stream = torch.cuda.Stream(device=dev)          # a real (non-null) stream shared by torch events and the library
    torch.cuda.set_stream(stream)
    blk.set_stream(stream.cuda_stream)
    L = q.load_library()
    overlap = not args.no_overlap
    if overlap:
        # streaming use: the loop / FEC tail of step k runs under the parallel stages of step k+1 (QRL_PARAM_OVERLAP_CALLS);
        # qrl_rx_join before the closing event puts every step's tail inside the timed region
        blk.set_overlap(True)

    def step_device():
        blk.work_device(X.data_ptr(), T, T)

    import ctypes as Ct
    sm_a, sm_b = Ct.c_int(), Ct.c_int()
    L.qrl_rx_sm_partition(blk._h, Ct.byref(sm_a), Ct.byref(sm_b))
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    n_bits = int(np.sum(blk.read_port_counts(2)))

    # ---- timed region: K steps, inputs resident in HBM (2.1 GB per step >> 126 MB L2: no flush needed)
    import ctypes as Ct
    L.qrl_rx_profile(blk._h, 1)
    launches0 = blk.launches
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(10.0)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.armed = True
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    blk.join()
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.armed = False
    if dist:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join()
    ms = e0.elapsed_time(e1)
    launches = blk.launches - launches0
    stage_ms = []
    for s in range(6):
        m, n = Ct.c_double(), Ct.c_long()
        L.qrl_rx_profile_read(blk._h, s, Ct.byref(m), Ct.byref(n))
        stage_ms.append((m.value, n.value))
    L.qrl_rx_profile(blk._h, 0)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * C * T * args.steps / (ms_max * 1e-3) / 1e6

    # ---- e2e: the reference-facing call with HOST buffers: H2D of the gr_complex slab + D2H of the decoded bits
    Xh = torch.empty((C, T), dtype=torch.complex64, pin_memory=True)
    Xh.copy_(X)
    bits_cap = int(blk.read_port_counts(2).max()) + 256
    out_bits = torch.empty((C, bits_cap), dtype=torch.uint8, pin_memory=True)
    out_cnt = np.zeros(C, np.int32)
    e2e_steps = max(2, min(args.steps, 5))

    def step_host():
        rc = L.qrl_rx_work(blk._h, Ct.c_void_p(Xh.data_ptr()), T, T, 0)
        assert rc == 0
        rc = L.qrl_rx_read_port(blk._h, 2, Ct.c_void_p(out_bits.data_ptr()), bits_cap, out_cnt.ctypes.data_as(Ct.c_void_p), 0)
        assert rc == 0

    step_host()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        step_host()
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    t2 = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * C * T * e2e_steps / (float(t2.item()) * 1e-3) / 1e6

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (stage-1 polyphase decimating FIR), live CUDA-event time
    peak, peak_src = peaks()
    fir_ms, fir_n = stage_ms[0]
    fir_avg_s = fir_ms / max(1, fir_n) * 1e-3
    launches_per_step = max(1, fir_n // args.steps)          # work() slices a step into several FIR launches
    alg_bytes = ALG_BYTES_PER_SAMPLE * C * T / launches_per_step
    achieved = alg_bytes / fir_avg_s / 1e9 if fir_avg_s > 0 else 0.0
    launches_per_step = max(1, stage_ms[0][1] // args.steps)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "fir_traffic_bytes.json")
    if os.path.exists(tp):      # ncu --set full capture (profiles/): DRAM bytes per input sample x samples per launch
        traffic = json.load(open(tp)).get("per_sample") * C * T / launches_per_step

    cpu_line = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, dt = cpu_chain_rate(12.0, cores, 1 << 20)
        cpu_line = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
                    "sample": "%d host threads x 2^20-sample chunks of the same 4FSK-2k-FM chain (CPU oracle port) for %.0f s" % (cores, dt)}

    line = {
        "metric": "RX IQ Msamples/s (batched channels)", "value": value, "unit": "Msamples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "channels_per_gpu": C, "samples_per_channel_per_step": T,
                   "l2": "inputs 2.1 GB/step > 126 MB L2, no flush", "parallelism": "channel-sharded x%d, no data-path collective" % world,
                   "calls": ("overlapped: tail of step k under step k+1 (QRL_PARAM_OVERLAP_CALLS), joined before the closing event" if overlap else "serialised"),
                   "decoded_bits_per_step": n_bits,
                   "sm_partition": {"loop_fec_sms": sm_a.value, "parallel_sms": sm_b.value}},
        "clocks": sampler.result(),
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(C * T * 8),
                "d2h_bytes_per_step": int(C * bits_cap + 4 * C), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "fir_decim_poly_kernel<50,9,8,128,8>", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": fir_avg_s * 1e3,
                     "launches_per_step": launches_per_step},
        "stage_ms_per_step": {n: (stage_ms[i][0] / args.steps) for i, n in enumerate(["fir_decim", "chan_filter", "demod_rrc", "symbol_sync", "viterbi", "soft_epilogue"])},
    }
    if cpu_line:
        line["cpu_baseline"] = cpu_line
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-overlap", action="store_true", help="serialise qrl_rx_work calls (no QRL_PARAM_OVERLAP_CALLS)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
