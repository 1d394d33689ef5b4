#!/usr/bin/env python
"""bench.py -- numrs_b200 headline benchmark (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--workload NAME]

Default workload = BASELINE.json configs[4] (the configuration the metric is quoted on):
rlft3 3-D real f64 512^3, one step = forward + inverse transform (isign=+1 then -1).
  N = 1 : single-GPU plan (device-resident, nrb_plan_exec).
  N > 1 : the volume is slab-decomposed across the N ranks (one process per GPU, launched by
          torchrun); one exchange per direction, fused into the FFT epilogue as stores into the peers'
          receive buffers over NVLink (--exchange nccl: NCCL all-to-all).  Strong scaling: the total work is fixed.
`value` = algorithmic HBM GB/s of the whole job: bytes every input element is read once and every
output element written once (2 151 677 952 B per direction at 512^3) / time.
`e2e` = the same metric through the reference-facing plugin call on HOST arrays: nrb_rlft3 / nrb_fourn (pinned host
volume, H2D + D2H inside the timed region); at N > 1 rank 0 makes that one call with the library spreading it over the
N GPUs itself (option num_devices, numrs_b200/csrc/multi.cpp) while the other ranks wait.

Other workloads (not driver defaults): rlft3_1024 (configs[4], 8 GPUs), fourn3d_512 (north_star's
3-D complex fourn), four1_batch (configs[1]), four1_1m (configs[0] batched x64), fourn2d (configs[2]), convlv, correl
(configs[3]).  The batch workloads shard BASELINE's fixed totals by batch across the ranks with no communication (strong scaling).

--impl reference times the reference's CPU path: the reference (Rust) cannot be compiled in
this image, so this is the oracle port (oracle/nr_oracle.c, which keeps the reference's
algorithm and loop structure) with all host threads (set explicitly: torchrun exports OMP_NUM_THREADS=1), on the
workload the line names (the full 512^3 volume for the default workload).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
NVLINK_GBS = 770.0          # measured peer copy per direction (profiling guide)


def rlft3_bytes(n1, n2, n3):
    return 8.0 * n1 * n2 * n3 * 2 + 16.0 * n1 * n2      # one direction: read + write data, speq


def rlft3_flops(n1, n2, n3):
    n = float(n1) * n2 * n3
    return 2.5 * n * np.log2(n)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "samples": len(sm),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


# =============================================================================== reference arm
def reference_workload(wl, O, cores):
    """The reference's CPU path (oracle port) on the workload `wl` names: returns (step, algorithmic bytes per step, sample
    description, whether the sample is the full workload)."""
    same_config = True
    if wl in RLFT3_DIMS:
        # the full volume the line's config names (a 512^3 forward + inverse takes seconds on the host cores)
        n1, n2, n3 = RLFT3_DIMS[wl]
        if wl == "rlft3_1024":      # a 1024^3 forward + inverse takes minutes on the host cores: bounded sample, and the line says so
            n1, n2, n3 = 512, 512, 512
            same_config = False
        x = O.fill_uniform(SEEDS[wl], 0, n1 * n2 * n3).reshape(n1, n2, n3)
        s = np.zeros((n1, 2 * n2))
        bytes_step = 2 * rlft3_bytes(n1, n2, n3)

        def step():
            O.rlft3(x, s, 1, mt=True)
            O.rlft3(x, s, -1, mt=True)
        sample = (f"rlft3 {n1}x{n2}x{n3} forward+inverse per step ({'the full volume' if same_config else '1/8 of the 1024^3 volume'}), "
                  f"oracle port with the reference's loop structure, {cores} OpenMP threads")
    elif wl == "fourn3d_512":
        n = 512
        x = O.fill_uniform(SEEDS[wl], 0, 2 * n ** 3)
        bytes_step = 2 * 32.0 * n ** 3

        def step():
            O.fourn(x, [n, n, n], 1, mt=True)
            O.fourn(x, [n, n, n], -1, mt=True)
        sample = f"fourn {n}^3 complex forward+inverse per step (the full volume), {cores} threads"
    elif wl in ("four1_batch", "four1_1m"):
        nn, cnt = (4096, 4096) if wl == "four1_batch" else (1 << 20, 64)
        arrs = [O.fill_uniform(SEEDS[wl], b * 2 * nn, 2 * nn) for b in range(cnt)]
        bytes_step = 2 * 32.0 * nn * cnt

        def step():
            O.fft_batch(arrs, 1, mt=True)
            O.fft_batch(arrs, -1, mt=True)
        sample = f"fft_batch {cnt} x four1({nn}) forward+inverse per step (the full batch), {cores} threads (one transform per thread, FFT_1.rs:186)"
    elif wl == "fourn2d":
        n = 8192
        x = O.fill_uniform(SEEDS[wl], 0, 2 * n * n)
        bytes_step = 2 * 32.0 * n * n

        def step():
            O.fourn(x, [n, n], 1, mt=True)
            O.fourn(x, [n, n], -1, mt=True)
        sample = f"fourn {n}x{n} forward+inverse per step (the full matrix), {cores} threads"
    else:
        # 256 signals of 2^22 points would take minutes per step on the host: a bounded sample of full-length signals,
        # one per thread as the reference's par_iter does (Convolve.rs:246, Correlation.rs:275); GB/s normalises the count
        n, m, cnt = 1 << 22, 4096, max(2, min(cores, 16))
        same_config = False
        sigs = [O.fill_uniform(1004, b * n, n) for b in range(cnt)]
        r = O.fill_uniform(1005, 0, m) / 64
        if wl == "convlv":
            bytes_step = 16.0 * n * cnt + 8.0 * n

            def step():
                O.convlv_batch(sigs, r, 1, mt=True)
        else:
            tm = [np.concatenate([O.fill_uniform(1005, 0, m), np.zeros(n - m)]) for _ in range(cnt)]
            bytes_step = 24.0 * n * cnt

            def step():
                O.correl_batch(sigs, tm, mt=True)
        sample = f"{wl}_batch {cnt} of the 256 signals, n=2^22, m=4096 per step, {cores} threads (one signal per thread)"
    return step, bytes_step, sample, same_config


def run_reference(args, rank, world):
    """Reference CPU path (oracle port, all host threads) on the workload the line names."""
    if rank != 0:
        return
    import oracle as O
    cores = O.use_all_cores()      # torchrun exports OMP_NUM_THREADS=1: state the thread count instead of inheriting it
    wl = args.workload
    step, bytes_step, sample, same_config = reference_workload(wl, O, cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = bytes_step * args.steps / dt / 1e9
    line = {"impl": "reference", "metric": METRIC[wl], "value": val, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": SCALING[wl], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_for(wl, args.gpus, None), sample_is_full_workload=same_config),
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


RLFT3_DIMS = {"rlft3_512": (512, 512, 512), "rlft3_1024": (1024, 1024, 1024)}
SEEDS = {"rlft3_512": 1006, "rlft3_1024": 1007, "fourn3d_512": 1008, "four1_batch": 1002, "four1_1m": 1001, "fourn2d": 1003,
         "convlv": 1004, "correl": 1004}
# BASELINE.json's fixed totals for the batch workloads (transforms / signals), sharded by batch over the ranks
BATCH_TOTAL = {"four1_batch": 4096, "four1_1m": 64, "convlv": 256, "correl": 256}
METRIC = {
    "rlft3_512": "rlft3 f64 512^3 forward+inverse algorithmic HBM GB/s",
    "rlft3_1024": "rlft3 f64 1024^3 forward+inverse algorithmic HBM GB/s",
    "fourn3d_512": "fourn f64 complex 512^3 forward+inverse algorithmic HBM GB/s",
    "four1_batch": "batched four1 f64 4096x4096 forward+inverse algorithmic HBM GB/s",
    "four1_1m": "batched four1 f64 64 x 2^20 forward+inverse algorithmic HBM GB/s",
    "fourn2d": "fourn f64 8192x8192 forward+inverse algorithmic HBM GB/s",
    "convlv": "convlv_batch f64 n=2^22 m=4096 algorithmic HBM GB/s",
    "correl": "correl_batch f64 n=2^22 algorithmic HBM GB/s",
}
SCALING = {"rlft3_512": "strong", "rlft3_1024": "strong", "fourn3d_512": "strong", "four1_batch": "strong", "four1_1m": "strong",
           "fourn2d": "weak", "convlv": "strong", "correl": "strong"}

EXCHANGE_TEXT = {
    "fused": "one exchange per direction fused into the FFT epilogue: stage-0 kernels store into the peers' receive buffers over NVLink (CUDA IPC peer memory), epoch-flag barrier kernels; no NCCL on the data path",
    "dma": "one exchange per direction: stage 0 writes a chunk-major send buffer, copy engines push the pieces into the peers' receive buffers over NVLink, per-chunk epoch flags",
    "nccl": "one NCCL all_to_all_single per direction",
}


def config_for(wl, gpus, exchange="fused"):
    def slab(what):
        if gpus == 1:
            return "single GPU"
        how = EXCHANGE_TEXT.get(exchange, "one exchange per direction (fused NVLink peer stores by default)")
        return f"{what} slab-decomposed over {gpus} GPUs (nn2-slabs in, nn1-slabs out), {how}"
    shard = (lambda tot, what: f"{tot} {what} sharded by batch over {gpus} GPU(s) ({tot // gpus} per GPU), no communication")
    base = {
        "rlft3_512": {"workload": "rlft3 3D real f64 512^3 (BASELINE configs[4]), step = forward + inverse",
                      "dims": [512, 512, 512], "parallelism": slab("volume")},
        "rlft3_1024": {"workload": "rlft3 3D real f64 1024^3 (BASELINE configs[4], the 8-GPU size), step = forward + inverse",
                       "dims": [1024, 1024, 1024], "parallelism": slab("volume")},
        "fourn3d_512": {"workload": "fourn 3D complex f64 512^3 (north_star: fourn/rlft3 512^3), step = forward + inverse",
                        "dims": [512, 512, 512], "parallelism": slab("complex volume")},
        "four1_batch": {"workload": "batched four1 f64 (BASELINE configs[1]): 4096 transforms of N=4096 in total, step = forward + inverse",
                        "parallelism": shard(4096, "transforms")},
        "four1_1m": {"workload": "batched four1 f64: 64 transforms of N=2^20 in total (BASELINE configs[0] batched), step = forward + inverse",
                     "parallelism": shard(64, "transforms")},
        "fourn2d": {"workload": "fourn 2D complex f64 8192x8192 (BASELINE configs[2]), step = forward + inverse",
                    "parallelism": "replicas only (one matrix per GPU)"},
        "convlv": {"workload": "convlv_batch f64 (BASELINE configs[3]): 256 signals of n=2^22 in total, m=4096, isign=+1",
                   "parallelism": shard(256, "signals")},
        "correl": {"workload": "correl_batch f64 (BASELINE configs[3]): 256 pairs of n=2^22 in total",
                   "parallelism": shard(256, "pairs")},
    }[wl]
    base["l2"] = "inputs larger than L2; every timed step works on a buffer not touched since the warm-up"
    return base


# =============================================================================== our arm
class Workload:
    """Device-resident workload: a pool of independent input buffers, one step = one buffer."""

    def __init__(self, lib, torch, name, pool, world=1, rank=0):
        self.lib, self.torch, self.name = lib, torch, name
        tot = BATCH_TOTAL.get(name, 1)
        if tot % world:
            raise SystemExit(f"bench.py: {tot} units do not divide over {world} GPUs")
        share = tot // world          # this rank's contiguous batch range [rank * share, (rank + 1) * share)
        import numrs_b200 as nb
        t = torch
        st = lambda: t.cuda.current_stream().cuda_stream  # noqa: E731
        self.stream = st
        f64 = dict(dtype=t.float64, device="cuda")
        if name == "four1_batch":
            nn, cnt = 4096, share
            self.plan = lib.plan_create(nb.KIND_FOUR1, [nn], batch=cnt)
            self.bufs = [t.empty(2 * nn * cnt, **f64) for _ in range(pool)]
            self.bytes_step = 2 * 32.0 * nn * cnt
            self.flops_step = 2 * 5.0 * nn * np.log2(nn) * cnt
            self.seed, self.scale = 1002, float(nn)
        elif name == "four1_1m":
            nn, cnt = 1 << 20, share
            self.plan = lib.plan_create(nb.KIND_FOUR1, [nn], batch=cnt)
            self.bufs = [t.empty(2 * nn * cnt, **f64) for _ in range(pool)]
            self.bytes_step = 2 * 32.0 * nn * cnt
            self.flops_step = 2 * 5.0 * nn * 20 * cnt
            self.seed, self.scale = 1001, float(nn)
        elif name == "fourn2d":
            n = 8192
            self.plan = lib.plan_create(nb.KIND_FOURN, [n, n], batch=1)
            self.bufs = [t.empty(2 * n * n, **f64) for _ in range(pool)]
            self.bytes_step = 2 * 32.0 * n * n
            self.flops_step = 2 * 5.0 * n * n * 26
            self.seed, self.scale = 1003, float(n * n)
        elif name in ("convlv", "correl"):
            n, m, cnt = 1 << 22, 4096, share
            self.n, self.cnt = n, cnt
            pool = min(pool, 2 if cnt > 64 else 4)
            self.bufs = [t.empty(n * cnt, **f64) for _ in range(pool)]
            self.out = t.empty(n * cnt, **f64)
            if name == "convlv":
                self.plan = lib.plan_create(nb.KIND_CONVLV, [n, m], batch=cnt)
                self.aux = t.empty(m, **f64)
                lib.fill_uniform_device(self.aux.data_ptr(), 1005, 0, m, st())
                self.aux /= 64.0
                self.bytes_step = 16.0 * n * cnt + 8.0 * n
                self.flops_step = cnt * (2 * 2.5 * n * 22 + 3.0 * n)
            else:
                self.plan = lib.plan_create(nb.KIND_CORREL, [n], batch=cnt)
                self.aux = t.zeros(n * cnt, **f64)
                tm = t.empty(m, **f64)
                lib.fill_uniform_device(tm.data_ptr(), 1005, 0, m, st())
                self.aux.view(cnt, n)[:, :m] = tm
                self.bytes_step = 24.0 * n * cnt
                self.flops_step = cnt * (3 * 2.5 * n * 22 + 3.0 * n)
            self.seed, self.scale = 1004, None
        else:
            raise SystemExit(f"unknown workload {name}")
        for i, b in enumerate(self.bufs):   # this rank's share of the seeded sequence
            lib.fill_uniform_device(b.data_ptr(), self.seed, rank * b.numel(), b.numel(), st())
        t.cuda.synchronize()
        self.launches_step = (self.plan.num_launches(1) + self.plan.num_launches(-1)) if self.scale else self.plan.num_launches(1)

    def step(self, i):
        b = self.bufs[i % len(self.bufs)]
        if self.scale:
            self.plan.exec(b.data_ptr(), isign=1, stream=self.stream())
            self.plan.exec(b.data_ptr(), isign=-1, stream=self.stream())
        else:
            self.plan.exec(b.data_ptr(), self.aux.data_ptr(), self.out.data_ptr(), isign=1, stream=self.stream())

    def profile(self, i):
        b = self.bufs[i % len(self.bufs)]
        if self.scale:
            return (self.plan.profile(b.data_ptr(), isign=1, stream=self.stream()) +
                    self.plan.profile(b.data_ptr(), isign=-1, stream=self.stream()))
        return self.plan.profile(b.data_ptr(), self.aux.data_ptr(), self.out.data_ptr(), isign=1, stream=self.stream())


def dominant_kernel(prof_runs):
    """prof_runs: list of [(name, bytes, ms)] -> (name, avg bytes/launch, avg ms/launch, share of step)."""
    agg = {}
    total = 0.0
    for run in prof_runs:
        for name, b, ms in run:
            a = agg.setdefault(name, [0.0, 0.0, 0])
            a[0] += b
            a[1] += ms
            a[2] += 1
            total += ms
    name = max(agg, key=lambda k: agg[k][1])
    b, ms, cnt = agg[name]
    table = {k: {"launches_per_step": v[2] / len(prof_runs), "ms_per_step": v[1] / len(prof_runs),
                 "algorithmic_GBps": (v[0] / v[1] / 1e6) if v[1] > 0 else None} for k, v in agg.items()}
    return name, b / cnt, ms / cnt, ms / total, table


def ncu_traffic(kernel_name):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return d.get(kernel_name, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def run_ours(args, rank, world, local_rank):
    import torch
    import numrs_b200 as nb
    lib = nb.lib()
    if not torch.cuda.is_available() or lib.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; numrs_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    lib.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, args.warmup
    wl = args.workload
    peak, peak_src = measured_peak()
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    f64 = dict(dtype=torch.float64, device="cuda")

    extra = {}
    is3d = wl in RLFT3_DIMS or wl == "fourn3d_512"
    if is3d:
        real = wl in RLFT3_DIMS
        n1, n2, n3 = RLFT3_DIMS[wl] if real else (512, 512, 512)
        vol = n1 * n2 * n3                               # points (real for rlft3, complex for fourn)
        vol_doubles = vol if real else 2 * vol
        bytes_step = 2 * (rlft3_bytes(n1, n2, n3) if real else 32.0 * vol)
        flops_step = 2 * (rlft3_flops(n1, n2, n3) if real else 5.0 * vol * np.log2(vol))
        rt_scale = (2.0 / vol) if real else (1.0 / vol)  # inverse(forward(x)) * rt_scale = x
        seed = SEEDS[wl]
        pool = max(2, min(K + W, 24 if vol_doubles <= (1 << 27) else 6))
        if world == 1:
            plan = lib.plan_create(nb.KIND_RLFT3 if real else nb.KIND_FOURN, [n1, n2, n3])
            bufs = [torch.empty(vol_doubles, **f64) for _ in range(pool)]
            speqs = [torch.empty(2 * n1 * n2 if real else 2, **f64) for _ in range(pool)]
            for b in bufs:
                lib.fill_uniform_device(b.data_ptr(), seed, 0, vol_doubles, st())
            launches_step = plan.num_launches(1) + plan.num_launches(-1)

            def step(i):
                b, s = bufs[i % pool], speqs[i % pool]
                plan.exec(b.data_ptr(), s.data_ptr(), isign=1, stream=st())
                plan.exec(b.data_ptr(), s.data_ptr(), isign=-1, stream=st())

            def profile(i):
                b, s = bufs[i % pool], speqs[i % pool]
                return (plan.profile(b.data_ptr(), s.data_ptr(), isign=1, stream=st()) +
                        plan.profile(b.data_ptr(), s.data_ptr(), isign=-1, stream=st()))

            def verify():
                ref = torch.empty(vol_doubles, **f64)
                lib.fill_uniform_device(ref.data_ptr(), seed, 0, vol_doubles, st())
                chk = ref.clone()
                sp = torch.empty(2 * n1 * n2 if real else 2, **f64)
                plan.exec(chk.data_ptr(), sp.data_ptr(), isign=1, stream=st())
                plan.exec(chk.data_ptr(), sp.data_ptr(), isign=-1, stream=st())
                chk.mul_(rt_scale)
                return float(torch.linalg.norm(chk - ref) / torch.linalg.norm(ref))
        else:
            from numrs_b200.dist_rlft3 import SlabRlft3
            G = world
            slab = SlabRlft3(lib, n1, n2, n3, mode=args.exchange, chunks=args.chunks, kind="rlft3" if real else "fourn")
            ld, sd, xd = slab.local_doubles, slab.speq_doubles, slab.xchg_doubles
            bufs = [torch.empty(ld, **f64) for _ in range(pool)]
            speq = torch.empty(sd, **f64) if real else None
            for i, b in enumerate(bufs):   # synthetic slabs: rank r's share of the seeded sequence, different data in every pool buffer
                lib.fill_uniform_device(b.data_ptr(), seed, (i * world + rank) * ld, ld, st())
            launches_step = slab.plan.num_launches(1) + slab.plan.num_launches(-1)
            extra["a2a_bytes_per_gpu_per_direction"] = slab.a2a_bytes_per_gpu()
            extra["exchange"] = EXCHANGE_TEXT[args.exchange] + (f" ({slab.chunks} z-chunk(s))" if slab.chunks > 1 else "")

            def one_direction(b, isign):
                slab.transform(b, speq, isign)

            def step(i):
                b = bufs[i % pool]
                one_direction(b, 1)
                one_direction(b, -1)

            profile = None

            def verify():
                ref = torch.empty(ld, **f64)
                lib.fill_uniform_device(ref.data_ptr(), seed, rank * ld, ld, st())
                chk = ref.clone()
                one_direction(chk, 1)
                one_direction(chk, -1)
                chk.mul_(rt_scale)
                num = (chk - ref).pow(2).sum()
                den = ref.pow(2).sum()
                t = torch.stack([num, den])
                dist.all_reduce(t)
                return float(torch.sqrt(t[0] / t[1]))
        shard = 1
    else:
        pool = max(2, min(K + W, 12))
        w = Workload(lib, torch, wl, pool, world if wl in BATCH_TOTAL else 1, rank if wl in BATCH_TOTAL else 0)
        bytes_step, flops_step, launches_step = w.bytes_step, w.flops_step, w.launches_step
        step, profile = w.step, w.profile
        shard = world          # per-rank bytes x ranks = the whole job (BASELINE totals are split over the ranks)

        def verify():
            if not w.scale:
                return None
            b = torch.empty_like(w.bufs[0])
            lib.fill_uniform_device(b.data_ptr(), w.seed, 0, b.numel(), st())
            ref = b.clone()
            w.plan.exec(b.data_ptr(), isign=1, stream=st())
            w.plan.exec(b.data_ptr(), isign=-1, stream=st())
            b.mul_(1.0 / w.scale)
            return float(torch.linalg.norm(b - ref) / torch.linalg.norm(ref))

    torch.cuda.synchronize()
    err = verify()

    # ---- timed region: W warm-up steps, then exactly K steps between barriers
    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        tms = torch.tensor([ms], device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms[0])
    value = bytes_step * shard * K / (ms * 1e-3) / 1e9

    # ---- per-kernel roofline (events around every launch, N = 1 plans)
    roof = None
    kern_table = None
    if profile is not None:
        runs = [profile(W + K + i) for i in range(min(K, 6))]
        name, b_l, ms_l, share, kern_table = dominant_kernel(runs[1:] if len(runs) > 1 else runs)
        ach = b_l / (ms_l * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": ncu_traffic(name), "peak_source": peak_src, "algorithmic_bytes_per_launch": b_l,
                "ms_per_launch": ms_l, "share_of_step": share}
    elif is3d:
        a2a = extra.get("a2a_bytes_per_gpu_per_direction", 0.0)
        roof = {"bound": "hbm", "kernel": "slab pipeline (z + x pass storing into the peers | flag barrier | y pass)", "achieved": value / world,
                "peak": peak, "unit": "GB/s", "frac": value / world / peak, "traffic": None, "peak_source": peak_src,
                "note": "per-GPU algorithmic GB/s of the whole step; the NVLink exchange is inside it",
                "nvlink_GBps_if_the_step_were_all_exchange": 2 * a2a / (ms / K * 1e-3) / 1e9}

    # ---- end to end through the host-slice C ABI (pinned host buffers, copies inside the timed region)
    e2e = None
    if is3d:
        Ke = max(1, min(K, 8 if vol_doubles <= (1 << 27) else 3))
        if dist:
            # the device-resident pools are no longer needed; rank 0's in-library multi-GPU call needs room on every device
            bufs.clear()
            if world > 1:
                slab.close()
            torch.cuda.empty_cache()
            dist.barrier()
        if rank == 0:
            # ONE process, the reference's call shape on whole host arrays; with N > 1 the library itself scatters the volume
            # over the N GPUs (multi.cpp: slab H2D over N PCIe links, fused NVLink exchange, gather)
            lib.set_option("num_devices", world)
            h = lib.pinned_empty(vol_doubles)
            src = torch.empty(vol_doubles, **f64)
            lib.fill_uniform_device(src.data_ptr(), seed, 0, vol_doubles, st())
            h[:] = src.cpu().numpy()
            del src
            torch.cuda.empty_cache()
            if real:
                hs = lib.pinned_empty(2 * n1 * n2)
                hv, hsv = h.reshape(n1, n2, n3), hs.reshape(n1, 2 * n2)

                def call(isign):
                    nb.rlft3(hv, hsv, n1, n2, n3, isign)
                api = "nrb_rlft3"
            else:
                def call(isign):
                    nb.fourn(h, [n1, n2, n3], 3, isign)
                api = "nrb_fourn (3-D)"
            before = lib.multi_device_calls(0)
            x0 = h[:4096].copy()
            call(1)
            call(-1)
            h *= rt_scale
            e2e_err = float(np.linalg.norm(h[:4096] - x0) / np.linalg.norm(x0))
            t0 = time.perf_counter()
            for _ in range(Ke):
                call(1)
                call(-1)
                h[0] *= 1.0     # the result is in host memory here
            dt = time.perf_counter() - t0
            io_bytes = 2 * 8 * vol_doubles + (2 * 16 * n1 * n2 if real else 0)
            e2e = {"value": bytes_step * Ke / dt / 1e9, "unit": "GB/s", "steps": Ke,
                   "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes, "ms_per_step": dt / Ke * 1e3,
                   "api": f"{api}(host arrays, pinned), forward + inverse, ONE call per direction from one process; the library spreads it over {lib.num_devices_in_use()} GPU(s)",
                   "devices_used": lib.num_devices_in_use(), "multi_device_calls": lib.multi_device_calls(0) - before,
                   "roundtrip_rel_l2_first_4096": e2e_err}
            lib.set_option("num_devices", 1)
        if dist:
            dist.barrier()
    else:
        e2e = run_e2e_batch(lib, nb, wl, world, rank, dist)

    # ---- CPU baseline (rank 0, N = 1): oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(wl)

    if rank == 0:
        line = {"metric": METRIC[wl], "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": SCALING[wl], "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_for(wl, world, args.exchange),
                "gflops": flops_step * shard * K / (ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": value / world / peak, "roundtrip_rel_l2": err,
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
                "gpu_launches": launches_step * K if launches_step else None, "kernels": kern_table}
        if world > 1 and is3d:
            line.update(extra)
        emit(line)
    if dist:
        dist.destroy_process_group()


def run_e2e_batch(lib, nb, wl, world, rank, dist):
    """End to end for the batch workloads: the reference-facing host-slice call on pinned host slices, from ONE process
    (rank 0); with N > 1 the library shards the batch over the N GPUs itself (option num_devices)."""
    import torch
    if dist:
        torch.cuda.empty_cache()
        dist.barrier()
    out = None
    if rank == 0:
        lib.set_option("num_devices", world)
        st = torch.cuda.current_stream().cuda_stream

        def pinned_fill(doubles, seed):
            h = lib.pinned_empty(doubles)
            step = 1 << 27
            for o in range(0, doubles, step):      # through a bounded device buffer: the host arrays may exceed what is free
                c = min(step, doubles - o)
                tmp = torch.empty(c, dtype=torch.float64, device="cuda")
                lib.fill_uniform_device(tmp.data_ptr(), seed, o, c, st)
                h[o:o + c] = tmp.cpu().numpy()
                del tmp
            return h
        before = lib.multi_device_calls(1)
        note = ""
        if wl in ("four1_batch", "four1_1m"):
            nn, cnt = (4096, 4096) if wl == "four1_batch" else (1 << 20, 64)
            h = pinned_fill(2 * nn * cnt, SEEDS[wl])
            table = lib.batch_table([h[2 * nn * b:2 * nn * (b + 1)] for b in range(cnt)])   # pointer / length tables built once

            def call():
                lib.check(lib.four1_batch_table(table, 1))
                lib.check(lib.four1_batch_table(table, -1))
            bytes_call, io = 2 * 32.0 * nn * cnt, 2 * 16 * nn * cnt
            api = "nrb_four1_batch (host slices, pinned), forward + inverse"
        elif wl == "fourn2d":
            n = 8192
            h = pinned_fill(2 * n * n, SEEDS[wl])

            def call():
                nb.fourn(h, [n, n], 2, 1)
                nb.fourn(h, [n, n], 2, -1)
            bytes_call, io = 2 * 32.0 * n * n, 2 * 16 * n * n
            api = "nrb_fourn (host array, pinned), forward + inverse"
        else:
            n, m, cnt = 1 << 22, 4096, 64      # 64 of the 256 signals: 2 GiB in + 2 GiB out of pinned host memory per call
            note = "; 64 of the 256 signals per call"
            h = pinned_fill(n * cnt, 1004)
            ho = lib.pinned_empty(n * cnt)
            sigs = [h[n * b:n * (b + 1)] for b in range(cnt)]
            outs = [ho[n * b:n * (b + 1)] for b in range(cnt)]
            if wl == "convlv":
                r = pinned_fill(m, 1005) / 64

                def call():
                    lib.check(lib.convlv_batch(sigs, r, 1, 0, outs)[0])
                bytes_call, io = 16.0 * n * cnt + 8.0 * n, 8 * n * cnt
                api = "nrb_convlv_batch (host slices, pinned)"
            else:
                h2 = lib.pinned_empty(n * cnt)
                h2[:] = 0.0
                tm = pinned_fill(m, 1005)
                for b in range(cnt):
                    h2[n * b:n * b + m] = tm
                seconds = [h2[n * b:n * (b + 1)] for b in range(cnt)]

                def call():
                    lib.check(lib.correl_batch(sigs, seconds, outs)[0])
                bytes_call, io = 24.0 * n * cnt, 16 * n * cnt
                api = "nrb_correl_batch (host slices, pinned)"
        call()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        dt = time.perf_counter() - t0
        out = {"value": bytes_call * reps / dt / 1e9, "unit": "GB/s", "steps": reps, "ms_per_step": dt / reps * 1e3,
               "h2d_bytes_per_step": io, "d2h_bytes_per_step": io if wl != "correl" else io // 2,
               "api": f"{api}, ONE call from one process; the library shards it over {lib.num_devices_in_use()} GPU(s){note}",
               "devices_used": lib.num_devices_in_use(), "multi_device_calls": lib.multi_device_calls(1) - before}
        lib.set_option("num_devices", 1)
    if dist:
        dist.barrier()
    return out


def cpu_baseline(wl):
    """The same CPU path as --impl reference, on the GPU box's host cores: about 10 - 30 s of work."""
    import oracle as O
    cores = O.use_all_cores()
    step, bytes_step, sample, same_config = reference_workload(wl, O, cores)
    t0 = time.perf_counter()
    reps = 0
    while reps < 1 or (time.perf_counter() - t0 < 10.0 and reps < 40):
        step()
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": bytes_step * reps / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x [{sample}]", "sample_is_full_workload": same_config}


_REAL_STDOUT = None


def emit(line):
    """The JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)          # libraries that print to fd 1 (NCCL version banner) must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rlft3_512", choices=sorted(METRIC))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", default="fused", choices=["fused", "dma", "nccl"], help="multi-GPU rlft3 exchange")
    ap.add_argument("--chunks", type=int, default=1, help="multi-GPU rlft3, fused exchange: z-chunks of the pipelined exchange (1 = off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
