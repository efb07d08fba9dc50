#!/usr/bin/env python
"""Benchmark of the hot path: NUTS sampling of poll_model_2020.stan.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --workload syn ...                        BASELINE config 5 on the streaming kernel family
  python bench.py --impl reference ...                      CPU arm (oracle restatement; see below)

Metric (BASELINE.json): leapfrog steps/sec (value) and ESS/sec (extra keys), whole job over N GPUs.
One STEP = one complete sampler run (warm-up + sampling) of `--chains` chains per GPU, fresh seed per step; chains are
sharded over ranks by global chain id (weak scaling: fixed chains per GPU) and the kept draws are exchanged with ONE NCCL
all-gather inside the timed region (its result is used: per-shard checksum in the line).
  --workload 2016 (default): the 2016 data list (S=51, T=254, 1258+361 polls, D=15098), 1024 chains x (500+500) -- the
      configuration the metric is quoted on (BASELINE configs 2-4); resident SMEM/TMEM kernel.
  --workload syn: SURVEY 8(d)'s synthetic S=256 x T=365 x N=50k list (BASELINE config 5), 148 chains x (40+10) per step;
      streaming kernel family.
`value` times potus_run with the data list already on the device (CUDA events, max over ranks); `e2e` times the public API
end to end from HOST buffers: named list in -> extract("predicted_score") of the kept draws (formed on the device, fp32 over
the bus) + summary() (election-day state table, national vote, electoral-college simulation, ESS / R-hat over EVERY sampling
iteration, reduced on the device), copies included.

The reference's own implementation of this path is rstan/CmdStan driven from R; neither exists in this image (nor can be
installed offline: /root/reference is an R project, pip has nothing to install), so the CPU arm times the fp64 C
restatement in oracle/ on the box's host cores and says so (`cpu_baseline.kind = "port"`): every chain starts from a
committed adapted oracle state and runs sampling-phase transitions, in the literal per-day mat-vec gradient form (the
reference's cost model; the reported value) and in the collapsed scan+GEMM form.  Its hand-coded gradient is far cheaper
than Stan's autodiff tape, so the GPU/CPU ratio understates the ratio against rstan.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Workloads.  "2016" = BASELINE.json configs[1..3] (the configuration the metric is quoted on; resident kernel);
# "syn" = configs[4], the synthetic S=256 x T=365 x N=50k stress problem of SURVEY.md 8(d) (streaming kernel family).
# algo_bytes: SURVEY.md 8(d): read q,p + write q,p once, fp32 state = 4*D*4.  dram_bytes: dram__bytes_read.sum +
# dram__bytes_write.sum per leapfrog of the kernel from the committed ncu --set full capture named in `ncu`.
WORKLOADS = {
    "2016": dict(D=15098, S=51, T=254, N=1619, kernel="potus_nuts_kernel", dram_bytes=(3.547104e9 + 6.019785e9) / 75480,
                 ncu="profiles/r02_resident_sampling_launch.txt: ncu capture of a sampling-phase launch, depth-8 trees", desc="poll_model_2020.stan, 2016 data list (S=51,T=254,N=1619,D=15098)",
                 data="2016 polls (reference data/all_polls.csv through the restated final_2016.R wrangling; committed fixture), random inits"),
    "syn": dict(D=144837, S=256, T=365, N=50000, kernel="potus_stream_kernel", dram_bytes=(846.641933e9 + 415.223650e9) / 150516,
                ncu="profiles/r02_stream_kernel.txt: ncu capture of a sampling-phase launch, depth-10 trees", desc="poll_model_2020.stan, synthetic list of SURVEY.md 8(d) (S=256,T=365,N=40000+10000,P=512,D=144837, seed 1843)",
                data="synthetic (generator of SURVEY.md 8(d), numpy PCG64 seed 1843), random inits"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def shard_chains(total_chains: int, world: int, rank: int):
    """Contiguous shard of global chain ids for `rank` (chain RNG streams are keyed by global id)."""
    base, rem = divmod(total_chains, world)
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, n


def allgather_draws(local, group=None):
    """The path's only collective: all-gather of the kept-draw buffer (same shape on every rank).
    Works on any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device_index: int):
        self.rows, self.proc, self.idx = [], None, device_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def usable_cores() -> int:
    """Host threads this process may actually use: affinity mask capped by the cgroup CPU quota
    (a 128-CPU box whose container is limited to N CPUs must not be reported as 128 cores)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per) + 0.5)))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return n


def ess_summary(pkg, monitor):
    """monitor: [chains, iter_sampling, S+1] on the logit scale -> Stan ESS over all chains."""
    d = pkg.diagnostics
    K = monitor.shape[-1]
    e = np.array([d.ess(monitor[:, :, k]) for k in range(K)])
    return float(np.nanmin(e)), float(np.nanmedian(e))


def run_cpu(args, data, pkg, reference_line: bool):
    """CPU arm: the oracle's C restatement (Stan-semantics NUTS, fp64, one chain per host thread).

    2016 workload: every chain starts from a committed ADAPTED oracle state (tests/golden/oracle_adapted_states_2016.npz:
    position after 500 warm-up iterations, its step size and inverse metric) and runs sampling-phase transitions -- the
    stationary depth-8 trajectories that make up the bulk of a run -- in BOTH gradient forms: the literal per-day mat-vec
    recurrence of poll_model_2020.stan:86 (the reference's cost model; the reported value) and the collapsed scan + GEMM form
    the GPU kernels use.  syn workload: bounded slice from random inits (a gradient costs ~50 ms on one core)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    orc.build()
    om = orc.OracleModel(data)
    cores = usable_cores()

    if args.workload != "2016":
        iters, depth = args.cpu_iters, 6

        def one(seed, literal=True):
            r = om.sample(chains=cores, iter_warmup=args.iter_warmup, iter_sampling=args.iter_sampling, seed=seed, threads=cores,
                          literal=literal, tree_mode=0, max_iters=iters, max_treedepth=depth)
            return int(r["n_leapfrog"].sum()), r["seconds"]
        sample = (f"{cores} chains x first {iters} warm-up iterations from random inits, max_treedepth {depth} (fp64 C restatement of Stan's NUTS; "
                  "per-day mat-vec gradient as in poll_model_2020.stan:86)")
        extra = {}
    else:
        from concurrent.futures import ThreadPoolExecutor
        with np.load(os.path.join(ROOT, "tests", "golden", "oracle_adapted_states_2016.npz")) as z:
            st = {k: z[k] for k in z.files}     # (eager: NpzFile is not thread-safe)
        nst, ntr = st["q"].shape[0], args.cpu_transitions

        def one(seed, literal=True):
            def chain(c):
                k = c % nst
                _, stats = om.transitions(st["q"][k].astype(np.float64), float(st["stepsize"][k]), st["inv_metric"][k].astype(np.float64),
                                          n_iter=ntr, seed=seed, chain=300000 + c, tree_mode=0, iter0=501, literal=literal)
                return int(stats[:, 4].sum())
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:       # ctypes releases the GIL: one chain per host thread
                lf = sum(ex.map(chain, range(cores)))
            return lf, time.perf_counter() - t0
        sample = (f"{cores} chains x {ntr} sampling-phase transitions each, started from committed adapted oracle states (post-warm-up position, "
                  "step size ~0.014, adapted diag metric; depth-8 trees); fp64 C restatement of Stan's NUTS, not rstan (absent from the image)")
        ora = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_posterior_2016.json")))
        extra = {"ess_per_sec_committed_run": {"min": min(ora["ess"]) / ora["seconds"], "median": float(np.median(ora["ess"])) / ora["seconds"],
                                               "what": f"tests/golden/oracle_posterior_2016.json: {ora['chains']} chains x (500+500), collapsed gradient, "
                                                       f"{ora['seconds']:.0f} s on 8 host threads of the authoring container (not this box)"}}

    if not reference_line:
        lf, secs = one(args.seed, True)
        lf2, secs2 = one(args.seed, False)
        return {"value": lf / secs, "unit": "leapfrog/s", "cores": cores, "kind": "port", "value_literal": lf / secs, "value_collapsed": lf2 / secs2,
                "sample": sample + f"; literal form {lf} leapfrogs in {secs:.1f} s, collapsed scan+GEMM form {lf2} in {secs2:.1f} s", **extra}
    for w in range(args.warmup):
        one(args.seed + 1000 + w)
    t_lf, t_s = 0, 0.0
    for k in range(args.steps):
        lf, secs = one(args.seed + k)
        t_lf += lf; t_s += secs
    v = t_lf / t_s
    lf2, secs2 = one(args.seed, False)
    cb = {"value": v, "unit": "leapfrog/s", "cores": cores, "kind": "port", "value_literal": v, "value_collapsed": lf2 / secs2,
          "sample": "each step: " + sample + " (value = literal form, the reference's cost model)", **extra}
    return {"metric": "leapfrog steps/sec", "value": v, "unit": "leapfrog/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_s / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": WORKLOADS[args.workload]["data"],
            "config": {"workload": WORKLOADS[args.workload]["desc"] + ", Stan-default NUTS", "chains": cores,
                       "iter_warmup": args.iter_warmup, "iter_sampling": args.iter_sampling,
                       "bounded_sample_per_step": sample},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "leapfrog/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="2016", choices=sorted(WORKLOADS), help="2016: BASELINE configs 2-4 (headline); syn: BASELINE config 5")
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (weak scaling); default 1024 (2016) / 148 (syn)")
    ap.add_argument("--iter-warmup", type=int, default=None, help="default 500 (2016) / 40 (syn: a step is a bounded slice of warm-up + sampling)")
    ap.add_argument("--iter-sampling", type=int, default=None, help="default 500 (2016) / 10 (syn)")
    ap.add_argument("--keep-per-chain", type=int, default=3, help="full draws kept per chain (1024x3 ~ the reference's 6x500)")
    ap.add_argument("--seed", type=int, default=1843)
    ap.add_argument("--cpu-iters", type=int, default=8, help="syn workload: bounded CPU sample, iterations per chain")
    ap.add_argument("--cpu-transitions", type=int, default=10, help="2016 workload: sampling-phase transitions per chain of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--warmup-scale", type=float, default=0.1,
                    help="untimed warm-up steps run the same chains for this fraction of the iterations (clock/cache warm-up)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    dflt = {"2016": (1024, 500, 500), "syn": (148, 40, 10)}[args.workload]
    args.chains = args.chains or dflt[0]
    args.iter_warmup = dflt[1] if args.iter_warmup is None else args.iter_warmup
    args.iter_sampling = dflt[2] if args.iter_sampling is None else args.iter_sampling
    algo_bytes = 4 * wl["D"] * 4
    gemm_flops = 4 * wl["S"] * wl["S"] * wl["T"]

    import potus_pkg
    pkg = potus_pkg.load()
    data = pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz")) if args.workload == "2016" else pkg.synthetic_datalist()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_cpu(args, data, pkg, True)))
        return

    import torch
    import torch.distributed as dist
    assert world == args.gpus, f"WORLD_SIZE={world} but --gpus {args.gpus}: launch with torch.distributed.run --nproc-per-node {args.gpus}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from us_potus_model_b200 import build as _b
    if local_rank == 0:
        _b.build()                      # (a no-op when the in-tree library is current; only one rank per node ever compiles)
    if world > 1:
        dist.barrier()
    model = pkg.cmdstan_model("poll_model_2020.stan")
    total_chains = args.chains * world
    off, n_local = shard_chains(total_chains, world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_run(seed, nw, ns, full_e2e):
        """One step.  Returns dict(device seconds, e2e seconds, leapfrogs, bytes, monitor, launches)."""
        t0 = time.perf_counter()
        fit = model.sample(data=data, seed=seed, chains=n_local, iter_warmup=nw, iter_sampling=ns,
                           keep_per_chain=args.keep_per_chain, device=local_rank, chain_id_offset=off)
        st = fit.stats
        # the path's one exchange: all-gather of the kept draws (device buffers, NVLink)
        ptr, n = fit.device_buffer(0)
        gather_bytes = 4 * n
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gather_ms = 0.0
        gather_checksum = None
        if world > 1 and n > 0:
            local = _wrap_device(ptr, n, dev)
            ev0.record()
            gathered = allgather_draws(local)
            ev1.record()
            torch.cuda.synchronize()
            gather_ms = ev0.elapsed_time(ev1)
            # use what arrived: every rank must hold every rank's draws (mu_b[,T] of the first kept draw of each shard)
            chk = gathered.view(world, -1)[:, :int(data["S"]) * int(data["T"])].double().sum(1)
            gather_checksum = [float(x) for x in chk.tolist()]
            del gathered
        out = {"dev_s": st["seconds_total"] + gather_ms * 1e-3, "warm_s": st["seconds_warmup"], "samp_s": st["seconds_sampling"],
               "lf": st["n_leapfrog_total"], "lf_samp": st["n_leapfrog_sampling"], "launches": st["gpu_launches"],
               "div": st["n_divergent_sampling"], "eps": st["mean_stepsize"], "depth": st["mean_treedepth"], "accept": st["mean_accept_stat"],
               "gather_ms": gather_ms, "gather_checksum": gather_checksum, "gather_bytes": gather_bytes}
        d2h = 0
        if full_e2e:
            # what a consumer of the reference's call gets back: the generated quantity the reports read (predicted_score of
            # the kept draws, formed on the device, fp32 over the bus) and the election-day summaries + ESS over ALL sampling
            # iterations, reduced on the device (potus_postprocess) -- the monitor table itself stays on the GPU
            ps = fit.extract("predicted_score"); d2h += ps.size * 4
            sm = fit.summary(ev=data.get("_ev_state"), ess=ns >= 4)
            d2h += (int(data["S"]) + 2) * 8 * 8 + ((int(data["S"]) + 1) * (n_local * 6 + ((n_local + 63) // 64) * ns) * 8 if ns >= 4 else 0)
            d2h += n_local * 72 + 64          # per-chain adaptation state + the device-reduced run statistics (inside potus_run)
            out["summary"] = sm
            out["pred_T"] = ps[:, -1, :]
        out["e2e_s"] = time.perf_counter() - t0
        out["d2h"] = d2h
        out["h2d"] = sum(np.asarray(v).nbytes for k, v in data.items() if not k.startswith("_"))
        fit.close()
        return out

    nw_w, ns_w = max(20, int(args.iter_warmup * args.warmup_scale)), max(5, int(args.iter_sampling * args.warmup_scale))
    for w in range(args.warmup):
        one_run(args.seed + 1000 + w, nw_w, ns_w, False)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    res = []
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        res.append(one_run(args.seed + k, args.iter_warmup, args.iter_sampling, True))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if rank == 0 else None

    dev_s = sum(r["dev_s"] for r in res)
    e2e_s = sum(r["e2e_s"] for r in res)
    lf = sum(r["lf"] for r in res)
    samp_s = sum(r["samp_s"] for r in res)
    lf_samp = sum(r["lf_samp"] for r in res)
    if world > 1:
        t = torch.tensor([dev_s, e2e_s, t_wall, samp_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, t_wall, samp_s = t.tolist()
        c = torch.tensor([lf, lf_samp], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        lf, lf_samp = c.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    value = lf / dev_s
    if args.iter_sampling >= 50:   # Stan ESS of the 52 monitored scalars, from the device post-processing of the last timed step
        ess_min, ess_med = float(np.nanmin(res[-1]["summary"]["ess"])), float(np.nanmedian(res[-1]["summary"]["ess"]))
    else:
        ess_min, ess_med = float("nan"), float("nan")
    run_s = res[-1]["dev_s"]
    samp_rate = lf_samp / samp_s if samp_s > 0 else float("nan")     # per-launch figure of the sampling-phase kernel
    per_gpu_rate = samp_rate / world
    line = {
        "metric": "leapfrog steps/sec", "value": value, "unit": "leapfrog/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 state / fp16x2-split tensor-core GEMM with f32 accumulate / f64 energy reductions",
        "data": wl["data"],
        "config": {"workload": wl["desc"] + f", Stan-default NUTS (diag_e, adapt_delta 0.8, max_treedepth 10), {args.iter_warmup}+{args.iter_sampling} iterations",
                   "chains_per_gpu": args.chains, "chains_total": total_chains, "parallelism": f"chains sharded over {world} GPU(s), one all-gather of draws",
                   "l2": ("per-step working set (per-chain state 268 MB + tree workspace 368 MB per GPU) exceeds the 126 MB L2; chain state itself is SMEM/TMEM-resident"
                          if args.workload == "2016" else
                          "per-step working set (tree workspace 26 MB per CTA x 148 + 2.3 MB per chain) exceeds the 126 MB L2; the state streams from HBM/L2 every leapfrog")},
        "ess_per_sec": {"min": ess_min / run_s * world, "median": ess_med / run_s * world, "quantities": "inv_logit-scale mu_b[,T] x51 + national",
                        "ess_min": ess_min, "ess_median": ess_med, "draws": int(n_local * args.iter_sampling),
                        "note": "ESS of this rank's chains over its run time (incl. warm-up), scaled by n_gpus"},
        "sampler": {"mean_stepsize": res[-1]["eps"], "mean_treedepth": res[-1]["depth"], "mean_accept_stat": res[-1]["accept"],
                    "divergent_sampling": int(res[-1]["div"]), "leapfrogs_per_step": lf / max(args.steps, 1)},
        "e2e": {"value": lf / e2e_s, "unit": "leapfrog/s", "h2d_bytes_per_step": int(res[-1]["h2d"]), "d2h_bytes_per_step": int(res[-1]["d2h"]),
                "api": "cmdstan_model().sample(data=<host named list>) + extract(predicted_score) + summary() (on-device state table / EV simulation / ESS)"},
        "gpu_launches": int(sum(r["launches"] for r in res)),
        "roofline": {"bound": "hbm", "achieved": per_gpu_rate * algo_bytes / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": per_gpu_rate * algo_bytes / 1e9 / peak,
                     "traffic": (wl["dram_bytes"] * (lf_samp / max(args.steps, 1) / world)) if wl["dram_bytes"] else None,
                     "traffic_note": f"bytes per sampling-phase launch = ncu DRAM bytes per leapfrog ({wl['ncu']}) x leapfrogs in the launch; "
                                     f"algorithmic bytes per launch = {algo_bytes} x leapfrogs",
                     "kernel": wl["kernel"] + " (sampling-phase launch)", "peak_source": peak_src,
                     "algorithmic_bytes_per_leapfrog": algo_bytes,
                     "tensor_frac_of_bf16_peak": per_gpu_rate * gemm_flops / 1700.3e12,
                     "tensor_frac_note": "4 S^2 T dense flops per leapfrog (SURVEY.md 8(d)) / measured bf16 burst peak 1700.3 TF/s; the kernels issue 3 fp16 "
                                         "products per GEMM (hi/lo split) on the lower-triangular half"},
        "clocks": clk, "wall_s_timed_region": t_wall,
        "allgather": {"ms_last_step": res[-1]["gather_ms"], "bytes_per_rank": int(res[-1].get("gather_bytes", 0)),
                      "checksum_per_shard_rank0": res[-1]["gather_checksum"]} if world > 1 else None,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = run_cpu(args, data, pkg, False)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _wrap_device(ptr: int, n: int, dev):
    """torch view of a device buffer owned by the sampler (no copy) via __cuda_array_interface__."""
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}

    return torch.as_tensor(_Arr(), device=dev)


if __name__ == "__main__":
    main()
