#!/usr/bin/env python
"""Per-kernel timing table (CUDA events around every launch, nrb_plan_profile) for a workload.
Usage: python tools/kernel_table.py [workload ...]   (NUMRS_B200_LIB selects a library build)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import numrs_b200 as nb  # noqa: E402

PEAK = 6650.0   # fallback of B200_PROFILING.md; MEASURED_PEAKS.json wins when present
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def table(prof_runs):
    agg = {}
    for run in prof_runs:
        for i, (name, b, ms) in enumerate(run):
            a = agg.setdefault(name, [0.0, 0.0, 0])
            a[0] += b
            a[1] += ms
            a[2] += 1
    rows = []
    for k, (b, ms, cnt) in agg.items():
        rows.append((ms / len(prof_runs), k, cnt / len(prof_runs), b / ms / 1e6, b / ms / 1e6 / PEAK))
    rows.sort(reverse=True)
    return rows


def main():
    lib = nb.lib()
    st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    f64 = dict(dtype=torch.float64, device="cuda")
    wls = sys.argv[1:] or ["rlft3_512"]
    out = {"lib": nb.LIB_PATH}
    for wl in wls:
        if wl.startswith("rlft3_"):
            n = int(wl.split("_")[1])
            plan = lib.plan_create(nb.KIND_RLFT3, [n, n, n])
            bufs = [torch.empty(n ** 3, **f64) for _ in range(4)]
            sp = torch.empty(2 * n * n, **f64)
            for b in bufs:
                lib.fill_uniform_device(b.data_ptr(), 1006, 0, b.numel(), st())
            run = lambda i: (plan.profile(bufs[i % 4].data_ptr(), sp.data_ptr(), isign=1, stream=st()) +  # noqa: E731
                             plan.profile(bufs[i % 4].data_ptr(), sp.data_ptr(), isign=-1, stream=st()))
            alg = 2 * (16.0 * n ** 3 + 16.0 * n * n)
        elif wl.startswith("four1_"):       # four1_<log2n>_<batch>
            _, lg, cnt = wl.split("_")
            nn, cnt = 1 << int(lg), int(cnt)
            plan = lib.plan_create(nb.KIND_FOUR1, [nn], batch=cnt)
            bufs = [torch.empty(2 * nn * cnt, **f64) for _ in range(4)]
            for b in bufs:
                lib.fill_uniform_device(b.data_ptr(), 1002, 0, b.numel(), st())
            run = lambda i: (plan.profile(bufs[i % 4].data_ptr(), isign=1, stream=st()) +  # noqa: E731
                             plan.profile(bufs[i % 4].data_ptr(), isign=-1, stream=st()))
            alg = 2 * 32.0 * nn * cnt
        elif wl.startswith("fourn2d_"):
            n = int(wl.split("_")[1])
            plan = lib.plan_create(nb.KIND_FOURN, [n, n])
            bufs = [torch.empty(2 * n * n, **f64) for _ in range(3)]
            for b in bufs:
                lib.fill_uniform_device(b.data_ptr(), 1003, 0, b.numel(), st())
            run = lambda i: (plan.profile(bufs[i % 3].data_ptr(), isign=1, stream=st()) +  # noqa: E731
                             plan.profile(bufs[i % 3].data_ptr(), isign=-1, stream=st()))
            alg = 2 * 32.0 * n * n
        elif wl.startswith("convlv_") or wl.startswith("correl_"):   # convlv_<log2n>_<batch>
            kind, lg, cnt = wl.split("_")
            n, cnt, m = 1 << int(lg), int(cnt), 4096
            a = torch.empty(n * cnt, **f64)
            lib.fill_uniform_device(a.data_ptr(), 1004, 0, a.numel(), st())
            o = torch.empty(n * cnt, **f64)
            if kind == "convlv":
                plan = lib.plan_create(nb.KIND_CONVLV, [n, m], batch=cnt)
                aux = torch.empty(m, **f64)
                lib.fill_uniform_device(aux.data_ptr(), 1005, 0, m, st())
                alg = 16.0 * n * cnt + 8.0 * n
            else:
                plan = lib.plan_create(nb.KIND_CORREL, [n], batch=cnt)
                aux = torch.zeros(n * cnt, **f64)
                alg = 24.0 * n * cnt
            run = lambda i: plan.profile(a.data_ptr(), aux.data_ptr(), o.data_ptr(), isign=1, stream=st())  # noqa: E731
        elif wl.split("_")[0] in ("twofft", "correlnorm", "correlnormfast", "autocorrel", "cosft1", "cosft2", "sinft"):
            kind, lg, cnt = wl.split("_")        # e.g. twofft_20_16, cosft1_16_1024
            n, cnt = 1 << int(lg), int(cnt)
            # three copies of the inputs, used in turn, so that a timed run never finds its lines in L2 (126 MB)
            nrot = 3 if (n + 2) * cnt * 8 < (512 << 20) else 1
            a_all = [torch.empty((n + 2) * cnt, **f64) for _ in range(nrot)]
            b_all = [torch.empty(n * cnt, **f64) for _ in range(nrot)]
            for t in a_all:
                lib.fill_uniform_device(t.data_ptr(), 1010, 0, t.numel(), st())
            for t in b_all:
                lib.fill_uniform_device(t.data_ptr(), 1011, 0, t.numel(), st())
            a, b2 = a_all[0], b_all[0]
            if kind == "twofft":
                plan = lib.plan_create(nb.KIND_TWOFFT, [n], batch=cnt)
                o = torch.empty(2 * (2 * n + 2) * cnt, **f64)
                alg = (16.0 + 32.0) * n * cnt           # two real inputs, two complex spectra
                run = lambda i: plan.profile(a_all[i % nrot].data_ptr(), b_all[i % nrot].data_ptr(), o.data_ptr(), isign=1, stream=st())  # noqa: E731
            elif kind in ("correlnorm", "correlnormfast", "autocorrel"):
                k = {"correlnorm": nb.KIND_CORREL_NORM, "correlnormfast": nb.KIND_CORREL_NORM_FAST,
                     "autocorrel": nb.KIND_AUTOCORREL_FAST}[kind]
                plan = lib.plan_create(k, [n], batch=cnt)
                o = torch.empty(4 * cnt + n * cnt, **f64)
                alg = (16.0 if kind == "autocorrel" else 24.0) * n * cnt
                run = lambda i: plan.profile(a_all[i % nrot].data_ptr(), b_all[i % nrot].data_ptr(), o.data_ptr(), isign=1, stream=st())  # noqa: E731
            else:
                k = {"cosft1": nb.KIND_COSFT1, "cosft2": nb.KIND_COSFT2, "sinft": nb.KIND_SINFT}[kind]
                plan = lib.plan_create(k, [n], batch=cnt)
                alg = 16.0 * n * cnt
                run = lambda i: plan.profile(a_all[i % nrot].data_ptr(), isign=1, stream=st())  # noqa: E731
        else:
            raise SystemExit(wl)
        torch.cuda.synchronize()
        runs = [run(i) for i in range(7)][1:]
        rows = table(runs)
        tot = sum(r[0] for r in rows)
        print(f"== {wl}  lib={os.path.basename(nb.LIB_PATH)}  step {tot:.3f} ms  algorithmic {alg / tot / 1e6:.0f} GB/s "
              f"({alg / tot / 1e6 / PEAK:.3f} of measured {PEAK:.0f})")
        for ms, name, cnt, gbs, frac in rows:
            print(f"   {ms:9.4f} ms  x{cnt:6.1f}  {gbs:8.0f} GB/s  {frac:5.3f}  {name}")
        out[wl] = {"step_ms": tot, "alg_GBps": alg / tot / 1e6, "kernels": [
            {"name": name, "ms_per_step": ms, "launches": cnt, "GBps": gbs, "frac": frac} for ms, name, cnt, gbs, frac in rows]}
        plan.destroy()
        torch.cuda.empty_cache()
    tag = os.path.splitext(os.path.basename(nb.LIB_PATH))[0]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"kernel_table_{tag}.json"), "a") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
