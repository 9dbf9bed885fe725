#!/usr/bin/env python
"""bench.py -- frames/s of the nnnoiseless per-frame denoise path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--streams B_per_gpu] [--frames T]

A "step" is one pass of the hot path over one batch: B streams per GPU, each advanced T consecutive
480-sample frames (T frame-steps of 5 kernels each).  Default workload = BASELINE.json configs[1]
(batch=4096 independent mono streams, built-in model, 1xB200) with T = 100 frames; at N
GPUs every rank owns its own B streams (weak scaling, no data-path collective; the model image is
broadcast once over NCCL).  Input = synthetic white+sine PCM-valued audio, resident in HBM before the
timed region; each step reads T*B*1920 B of input (786 MB by default, > the 126 MB L2).

--impl reference times the reference's CPU implementation of the same path (the C restatement in
oracle/ -- the Rust crate cannot be built in this image) with all host threads, on a bounded sample of
the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME = 480
# SURVEY 8(d) / BASELINE.md section 3: algorithmic bytes per stream-frame at T = 1 frame per launch
BYTES_IO = 3844            # 1920 in + 1920 out + 4 vad
BYTES_STATE = 13016        # state round trip (8308 read + 4708 written)
BYTES_PER_FRAME = BYTES_IO + BYTES_STATE  # 16860
# algorithmic bytes of each kernel of the 5-kernel pipeline taken alone (its own compulsory I/O per frame)
KERNEL_BYTES = {
    "hp_filter": 1920 + 8 + 1920 + 8,
    "pitch": 1728 * 4 + 8 + 4 + 8,
    "analysis": 960 * 4 + 960 * 4 + 4 + 704 + 4 + 3848 + 3200 + 3 * 88 + 168 + 4 + 88 + 4,
    "rnn": 168 + 672 + 4 + 672 + 88 + 4,
    "synthesis": 3848 + 3200 + 3 * 88 + 88 + 88 + 88 + 1920 + 1920 + 1920 + 4,
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); smax.append(float(c[2]))
                except ValueError:
                    continue
                for n, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_on_device(torch, B, T, device, seed):
    """[T][B][480] float32 on `device`: clamp(round(A sin(2 pi f n/48000 + phi) + sigma N(0,1))), per-stream
    f in [100,4000] Hz log-uniform, A in [1000,12000], sigma in [100,3000]."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    r = torch.rand(4, B, generator=g, device=device, dtype=torch.float64)
    f = 100.0 * torch.pow(torch.tensor(40.0, dtype=torch.float64, device=device), r[0])
    a = (1000.0 + 11000.0 * r[1]).float()
    sg = (100.0 + 2900.0 * r[2]).float()
    ph = 2 * np.pi * r[3]
    x = torch.empty(T, B, FRAME, device=device, dtype=torch.float32)
    n = torch.arange(FRAME, device=device, dtype=torch.float64)
    for t in range(T):
        arg = 2 * np.pi * f[:, None] * (n[None, :] + t * FRAME) / 48000.0 + ph[:, None]
        arg = torch.remainder(arg, 2 * np.pi).float()
        v = a[:, None] * torch.sin(arg) + sg[:, None] * torch.randn(B, FRAME, generator=g, device=device)
        x[t] = torch.clamp(torch.round(v), -32768.0, 32767.0)
    return x


# BASELINE.json "metric": "48kHz mono frames/sec (480-sample) at 1/2/4/8 B200; %HBM roofline; vs Rust CPU" -- the
# throughput part is the line's value, the other two parts are the `roofline` and `cpu_baseline` objects of the line.
METRIC = "48kHz mono frames/sec (480-sample)"


def workload_name(streams_per_gpu, frames):
    """config.workload, identical for both arms (BASELINE.json configs[1] when streams_per_gpu = 4096)."""
    return "configs[1]: batch=%d independent mono streams per GPU x %d frames per step, built-in model" % (streams_per_gpu, frames)


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arm must not obey that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline_run(x_bt, threads=None):
    """Times the oracle (C restatement of the reference) on [n][T][480] host samples; returns frames/s, threads."""
    import oracle
    if threads is None:
        threads = host_threads()
    import nnnoiseless_b200 as nb
    with open(nb.BUILTIN_WEIGHTS_PATH, "rb") as f:
        m = oracle.Model(f.read())
    r = oracle.run_batch(m, x_bt, n_threads=threads, want_out=True, want_taps=False)
    return x_bt.shape[0] * x_bt.shape[1] / r["seconds"], r["threads"]


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    T = args.frames
    n = min(args.streams, max(cores * 8, 8))
    from nnnoiseless_b200.synth import synth_streams  # numpy generator, same signal family as the GPU arm
    x = synth_streams(min(n, 64), T, seed=1234).reshape(-1, T, FRAME)
    reps = (n + x.shape[0] - 1) // x.shape[0]
    x = np.concatenate([x] * reps)[:n]
    for _ in range(args.warmup):
        cpu_baseline_run(x[: max(cores, 1)])
    t0 = time.perf_counter()
    frames = 0
    threads = 1
    for _ in range(args.steps):
        fps, threads = cpu_baseline_run(x)
        frames += x.shape[0] * T
    dt = time.perf_counter() - t0
    value = frames / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.streams, T), "streams_per_gpu": args.streams, "frames_per_step": T},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d streams x %d frames per step, oracle/nno_oracle.c (C restatement; no rustc in image), "
                                   "OpenMP one stream per thread" % (n, T)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    import nnnoiseless_b200 as nb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout; stdout carries exactly one JSON line
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # model image: rank 0 owns it, everybody else receives it over NCCL (the only collective of the path)
    if world > 1:
        n = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            img = nb.RnnModel().to_bytes()
            n[0] = len(img)
        dist.broadcast(n, 0)
        buf = torch.zeros(int(n.item()), dtype=torch.uint8, device=dev)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(img), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        model = nb.RnnModel.from_bytes(buf.cpu().numpy().tobytes())
        assert model is not None
    else:
        model = nb.RnnModel()

    B, T = args.streams, args.frames
    batch = nb.DenoiseBatch(B, model, device=local)
    x = synth_on_device(torch, B, T, dev, seed=1234 + rank)
    out = torch.empty_like(x)
    vad = torch.empty(T, B, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step():
        batch.process_device(out.data_ptr(), x.data_ptr(), vad.data_ptr(), T, stream_stride=FRAME, frame_stride=B * FRAME,
                             cuda_stream=sp)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    if rank == 0:
        # nvidia-smi needs a moment to start: keep the GPU under the same load until the first sample has arrived
        t_wait = time.time()
        while sampler.proc is not None and os.path.getsize(sampler.path) == 0 and time.time() - t_wait < 5.0:
            step()
            torch.cuda.synchronize()
    barrier()
    l0 = nb.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = nb.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    frames_total = world * B * T * args.steps
    value = frames_total / (ms_total * 1e-3)

    # ---- per-kernel durations (CUDA events between the kernels, same stream, same workload) ----
    kms = {}
    nprof = min(T, 20)
    for t in range(nprof):
        d = batch.profile_step(out[t].data_ptr(), x[t].data_ptr(), vad[t].data_ptr(), FRAME, sp)
        for k, v in d.items():
            kms.setdefault(k, []).append(v)
    kavg = {k: float(np.mean(v)) for k, v in kms.items()}
    step_ms = sum(kavg.values())
    dom = max(kavg, key=kavg.get)
    peak, peak_src = load_peaks()
    # one launch set = one frame of the rank's B streams; its device time inside the timed region (stages of
    # consecutive frames overlap) is ms_total / (frames per rank) -- all ranks run the same schedule
    launch_ms = ms_total / (T * args.steps)
    path_gbs = BYTES_PER_FRAME * B / (launch_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
        "achieved": path_gbs, "frac": path_gbs / peak, "frame_step_ms_pipelined": launch_ms,
        # DRAM bytes per frame-step measured with ncu --set full (profiles/r01_v6_ncu_raw_B65536.csv: dram__bytes_read +
        # dram__bytes_write of the five kernels = 2.40 GB at B = 65,536 = 36,600 B per stream-frame), scaled to this B
        "traffic": 36600.0 * B, "traffic_unit": "bytes per frame-step (ncu, scaled from B=65536)",
        "dominant_kernel_traffic": {"kernel": "pitch", "bytes_per_launch": 6982.0 * B, "algorithmic_bytes_per_launch": KERNEL_BYTES["pitch"] * B},
        "definition": "16,860 algorithmic B/frame (T=1: 3,844 I/O + 13,016 state round trip, SURVEY 8(d)) x %d frames per "
                      "frame-step / CUDA-event time per frame-step inside the timed region (the five kernels of a frame-step; "
                      "kernels of up to 4 consecutive frames overlap on separate streams)" % B,
        "frame_step_ms_serial_sum": step_ms, "dominant_kernel": dom,
        "kernels": {k: {"ms": v, "share": v / step_ms, "own_bytes_per_frame": KERNEL_BYTES.get(k),
                        "own_gbs": (KERNEL_BYTES.get(k, 0) * B / (v * 1e-3) / 1e9) if v > 0 else None}
                    for k, v in kavg.items()},
        "io_only_frac": (BYTES_IO * B / (launch_ms * 1e-3) / 1e9) / peak,
        "compute_note": "path is FP32-issue/latency bound, not HBM bound (SURVEY 8(d)); frac is reported against HBM as asked",
    }

    # ---- e2e: same metric through the public host-buffer API (pinned host memory, copies inside the timed region) ----
    Te = min(T, args.e2e_frames) if args.e2e_frames > 0 else T
    while Te > 1 and Te * B * FRAME * 4 > (512 << 20):  # keep each pinned staging buffer under 512 MiB
        Te //= 2
    hx = torch.empty(Te, B, FRAME, dtype=torch.float32).pin_memory()
    hx.copy_(x[:Te].cpu())
    ho = torch.empty(Te, B, FRAME, dtype=torch.float32).pin_memory()
    hv = torch.empty(Te, B, dtype=torch.float32).pin_memory()
    L = nb.lib()
    import ctypes as C

    def e2e_step():
        rc = L.rnnoise_batch_process_host(batch._h, C.c_void_p(ho.data_ptr()), C.c_void_p(hx.data_ptr()),
                                          C.c_void_p(hv.data_ptr()), Te)
        assert rc == 0, nb.last_error()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    ne = max(1, min(args.steps, 5))
    for _ in range(ne):
        e2e_step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * Te * ne / float(dt.item())
    e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": Te * B * FRAME * 4,
           "d2h_bytes_per_step": Te * B * (FRAME * 4 + 4), "frames_per_call": Te,
           "api": "rnnoise_batch_process_host (pinned host buffers, H2D + 5 kernels/frame + D2H, synchronous)"}
    # same through the 16-bit PCM entry point (int16 in/out, conversion fused into the kernels): half the PCIe bytes
    hx16 = torch.empty(Te, B, FRAME, dtype=torch.int16).pin_memory()
    hx16.copy_(x[:Te].to(torch.int16).cpu())
    ho16 = torch.empty(Te, B, FRAME, dtype=torch.int16).pin_memory()

    def e2e16_step():
        rc = L.rnnoise_batch_process_pcm16_host(batch._h, C.c_void_p(ho16.data_ptr()), C.c_void_p(hx16.data_ptr()),
                                                C.c_void_p(hv.data_ptr()), Te)
        assert rc == 0, nb.last_error()

    e2e16_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(ne):
        e2e16_step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e["pcm16"] = {"value": world * B * Te * ne / float(dt.item()), "unit": "frames/s", "h2d_bytes_per_step": Te * B * FRAME * 2,
                    "d2h_bytes_per_step": Te * B * (FRAME * 2 + 4), "api": "rnnoise_batch_process_pcm16_host"}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        n = min(B, max(8, 16 * cores))
        xs = x[:, :n].permute(1, 0, 2).contiguous().cpu().numpy()  # [n][T][480]
        cpu_baseline_run(xs[: max(1, min(n, cores))])  # warm-up (tables, page faults)
        fps, threads = cpu_baseline_run(xs)
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "first %d streams x %d frames of the GPU workload; oracle/nno_oracle.c (C restatement of the "
                         "reference; Rust toolchain absent), -O3 -march=native -ffp-contract=off, OpenMP one stream per thread"
                         % (n, T)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(B, T),
                       "streams_per_gpu": B, "frames_per_step": T, "parallelism": "streams sharded x%d, no data-path collective" % world,
                       "l2_policy": "inputs larger than L2: each step streams %.0f MB in + %.0f MB out through HBM"
                                    % (T * B * 1920 / 1e6, T * B * 1920 / 1e6)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU (configs[1] = 4096)")
    ap.add_argument("--frames", type=int, default=100, help="frames per stream per step")
    ap.add_argument("--e2e-frames", type=int, default=0, help="frames per host-API call in the e2e leg (0 = --frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
