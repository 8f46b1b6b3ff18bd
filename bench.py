#!/usr/bin/env python
"""Benchmark of the hot path on B200: influenza_1718 tau-leap trajectories/s (BASELINE.json metric, config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                                 (CPU arm: the oracle port on all host cores)

One "step" = one pass of the hot path over one batch: N_TRAJ = 10^6 influenza trajectories (28 compartments,
tau = 0.5, 2017-12-01 -> 2018-05-11 = 322 steps, 162 output days) with per-trajectory draws of beta and f_ud,
int32 output [28][162][N_TRAJ] (18.1 GB per step -- far larger than the 126 MB L2, so no step can reuse cached
data; the drawn inputs are re-uploaded for e2e).  Each rank runs its own N_TRAJ (weak scaling, no data-path
collective: trajectories are independent); Philox streams are keyed by global trajectory id.

JSON keys follow the driver contract; `value` = device-resident throughput (CUDA events, max over ranks),
`e2e` = through `pysodm_b200`'s host pipeline (pinned host draws in, full ensemble in the reference's
(draws, time, age) layout out, H2D/D2H inside the timed region), `roofline` / `cpu_baseline` as documented in
DESIGN.md.
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
sys.path.insert(0, ROOT)

N_TRAJ = 1_000_000
N_OUT = 162
N_STEPS = 322
TAU = 0.5
N_STATE = 28
BYTES_PER_TRAJ = N_STATE * N_OUT * 4 + 5 * 8 + 4      # int32 output + 5 drawn FP64 parameters + status word
VARIATES_PER_TRAJ = N_STEPS * 24
METRIC = "influenza_1718 tau-leap trajectories/sec"
UNIT = "trajectories/s"


def synthetic_draws(n, seed):
    """beta_i = 0.0174 (1 + 0.05 z), f_ud_i = [.01,.64,.905,.60] * U(0.9, 1.1), clipped."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal(n)
    u = rng.uniform(0.9, 1.1, size=(4, n))
    beta = np.clip(0.0174 * (1 + 0.05 * z), 1e-6, 0.08)
    f_ud = np.clip(np.array([0.01, 0.64, 0.905, 0.60])[:, None] * u, 1e-3, 1 - 1e-3)
    return np.ascontiguousarray(np.vstack([beta[None, :], f_ud]))          # [5, n]: rows = param elements 1, 4..7


VAR_INDEX = (1, 4, 5, 6, 7)


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk = float(f[1])
                smax = float(f[2])
            except ValueError:
                continue
            if t_begin - 0.05 <= ts <= t_end + 0.05:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:      # timed region shorter than the sampling period: use every sample taken
            sm = [float(l.split(",")[1]) for _, l in self.lines if l.count(",") >= 8 and l.split(",")[1].strip().replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (NumPy restatement, batched over trajectories) on all host cores
# ------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n = args
    from oracle import jump, models
    from pysodm_b200 import tutorial_data as td
    draws = synthetic_draws(n, seed)
    par = dict(td.INFLUENZA_PARAMETERS)
    par.update(beta=draws[0], f_ud=np.ascontiguousarray(draws[1:].T))
    par.pop("N")
    tdpf = {"N": (td.CONTACT_TABLE, td.influenza_schedule())}
    y = jump.tau_leap_ensemble(models.InfluenzaJump, td.influenza_flat_y0(), par, 0, N_OUT - 1, TAU, n, tdpf=tdpf,
                               seed=seed, chunk=1024)
    return float(y[:, 24:28, -1].mean())


def cpu_reference_rate(per_core=1024, rounds=1, pool=None, cores=None):
    """Throughput of the oracle port: `cores` processes x `per_core` trajectories per round."""
    import multiprocessing as mp
    cores = cores or os.cpu_count()
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores)
    t = time.perf_counter()
    for r in range(rounds):
        pool.map(_cpu_worker, [(1000 * r + c, per_core) for c in range(cores)])
    dt = time.perf_counter() - t
    if own:
        pool.close()
    return cores * per_core * rounds / dt, dt, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = os.cpu_count()
    per_core = 512
    pool = mp.get_context("fork").Pool(cores)
    for _ in range(args.warmup):
        cpu_reference_rate(per_core=64, pool=pool, cores=cores)
    t = time.perf_counter()
    for k in range(args.steps):
        pool.map(_cpu_worker, [(7000 * k + c, per_core) for c in range(cores)])
    dt = time.perf_counter() - t
    pool.close()
    n = cores * per_core * args.steps
    value = n / dt
    sample = (f"{cores} processes x {per_core} trajectories x {args.steps} steps of the same workload "
              f"(oracle/jump.py: NumPy tau-leap vectorised over trajectories, reference law incl. the stay-probability "
              f"quirk, legacy NumPy binomial); the unmodified reference itself runs ~5.6 traj/s/core "
              f"(BASELINE.md) because it loops in Python per trajectory")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/i64", "data": "synthetic",
            "config": {"workload": "influenza_1718 tau-leap tau=0.5, 2017-12-01..2018-05-11, per-trajectory beta/f_ud draws",
                       "n_traj_per_step": cores * per_core},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pysodm_b200 import engine, pipeline
    from pysodm_b200 import tutorial_data as td

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N with N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    dev = engine.device(local)
    all_cpus = os.sched_getaffinity(0)
    if not args.no_affinity:
        # one process per GPU: run (and first-touch the pinned result buffers) on the CPU cores of the GPU's own NUMA node
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device()))
        except Exception as exc:  # noqa: BLE001  (affinity is an optimisation, never a requirement)
            print(f"bench: CPU affinity not set ({exc})", file=sys.stderr)
    if world > 1:
        # NCCL prints its version banner to STDOUT when the communicator is created; stdout carries exactly one JSON line,
        # so file descriptor 1 points at stderr until the first collective has gone through
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    n = args.n_traj
    y0 = torch.from_numpy(td.influenza_flat_y0()).to(dev)
    par = td.influenza_flat_par()
    table = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).to(dev)
    index = torch.from_numpy(td.influenza_schedule()).to(dev)
    draws_host = torch.from_numpy(synthetic_draws(n, 1718 + rank)).pin_memory()          # [5, n] pinned
    draws_dev = draws_host.to(dev)
    out = torch.empty((N_STATE, N_OUT, n), dtype=torch.int32, device=dev)
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    var = (VAR_INDEX, draws_dev)

    def step(seed):
        engine.jump_run("influenza_1718_jump", y0, par, n, 0, N_OUT, TAU, var=var, tdpf=(table, index), seed=seed,
                        traj_id0=rank * n, out_dtype=torch.int32, out=out, status=status, dev=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput -----------------------------------------------------------------------
    for w in range(args.warmup):
        step(100 + w)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_begin = time.time()
    ev[0].record()
    for k in range(args.steps):
        step(1000 + k)
        ev[k + 1].record()
    barrier()
    t_end = time.time()
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    if int(status.abs().sum()) != 0:
        raise SystemExit("bench: trajectories reported a non-zero status")
    checksum = int(out[24:28, -1].to(torch.int64).sum())

    # ---- end to end through the host pipeline -------------------------------------------------------------
    import psutil
    n_e2e = args.e2e_traj or n
    need = n_e2e * N_STATE * N_OUT * 4
    avail = psutil.virtual_memory().available
    per_rank_budget = 0.5 * avail / max(1, world)
    if need > per_rank_budget:
        n_e2e = max(65536, int(per_rank_budget // (N_STATE * N_OUT * 4)) // 65536 * 65536)
    del out
    torch.cuda.empty_cache()
    variables, n_row = pipeline.layout_from_shapes({s: (4,) for s in td.INFLUENZA_STATES})
    host_out = pipeline.alloc_host_result(variables, n_e2e, N_OUT, torch.int32)
    pipe = pipeline.HostPipeline(variables, n_row, N_OUT, args.chunk, torch.int32, torch.int32, dev=dev)
    y0_host = torch.from_numpy(td.influenza_flat_y0()).pin_memory()
    table_host = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).pin_memory()
    index_host = torch.from_numpy(td.influenza_schedule()).pin_memory()
    h2d_bytes = draws_host[:, :n_e2e].numel() * 8 + y0_host.numel() * 8 + table_host.numel() * 8 + index_host.numel() * 4
    d2h_bytes = sum(t.numel() * 4 for t in host_out.values())

    def e2e_step(seed):
        # every input of the step starts in (pinned) host memory; the full ensemble ends in host memory
        y0_d = y0_host.to(dev, non_blocking=True)
        tab_d = table_host.to(dev, non_blocking=True)
        idx_d = index_host.to(dev, non_blocking=True)

        def launch(lo, m, buf, stream):
            chunk_draws = draws_host[:, lo:lo + m].to(dev, non_blocking=True)        # strided pinned -> device
            engine.jump_run("influenza_1718_jump", y0_d, par, m, 0, N_OUT, TAU, var=(VAR_INDEX, chunk_draws), tdpf=(tab_d, idx_d), seed=seed,
                            traj_id0=rank * n + lo, out_dtype=torch.int32, out=buf, dev=dev, stream=stream)
        pipe.run(n_e2e, launch, host_out)
        pipe.synchronize()

    e2e_warm = min(args.warmup, 2)
    for w in range(e2e_warm):
        e2e_step(100 + w)
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(1000 + k)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_launches = pipe.kernel_launches
    if n_e2e == n:
        e2e_checksum = int(host_out["Im_inc"][:, -1, :].to(torch.int64).sum())
        last_seed_match = (e2e_steps == args.steps)
        if last_seed_match and e2e_checksum != checksum:
            raise SystemExit("bench: host pipeline result differs from the device-resident run")

    # ---- the tutorial workflow, reduced on the device (N = 1 only) ------------------------------------------
    # sim -> negative-binomial observation noise -> mean and 95 % band over draws (calibration.py:231,245-246), with
    # only the drawn parameters going in and the (time, age) summaries coming out over PCIe
    workflow = None
    if not args.no_workflow:
        from pysodm_b200 import sharding
        del host_out, pipe
        torch.cuda.empty_cache()
        ens = torch.empty((N_STATE, N_OUT, n), dtype=torch.int32, device=dev)
        noisy = torch.empty_like(ens)
        n_all = n * world                                       # the ensemble of ALL ranks: reductions cross the GPUs
        lo_r, hi_r = int(np.floor((n_all - 1) * 0.025)), int(np.floor((n_all - 1) * 0.975))
        ranks = [lo_r, min(lo_r + 1, n_all - 1), hi_r, min(hi_r + 1, n_all - 1)]
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

        def workflow_step(seed):
            d = draws_host.to(dev, non_blocking=True)
            marks[0].record()
            engine.jump_run("influenza_1718_jump", y0, par, n, 0, N_OUT, TAU, var=(VAR_INDEX, d), tdpf=(table, index),
                            seed=seed, traj_id0=rank * n, out_dtype=torch.int32, out=ens, status=status, dev=dev)
            marks[1].record()
            engine.noise("negative_binomial", ens, 0.03, seed=seed, traj_id0=rank * n, out=noisy, dev=dev)
            marks[2].record()
            mean = sharding.sum_sharded(engine.draws_sum(noisy, dev=dev))                 # N > 1: all-reduce of the sums
            marks[3].record()
            band = sharding.order_stats_sharded(noisy, ranks)        # N > 1: radix select in phases, histograms all-reduced
            marks[4].record()
            return mean.cpu(), band.cpu()                      # the step's result reaches the host

        workflow_step(50)
        barrier()
        wf_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for k in range(wf_steps):
            mean_h, band_h = workflow_step(2000 + k)
        barrier()
        wf_s = time.perf_counter() - t0
        stages = [marks[i].elapsed_time(marks[i + 1]) for i in range(4)]
        wf_t = torch.tensor([wf_s] + stages, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(wf_t, op=dist.ReduceOp.MAX)
        wf_s, stages = float(wf_t[0]), [float(x) for x in wf_t[1:]]
        workflow = {"value": n_all * wf_steps / wf_s, "unit": UNIT, "ms_per_step": wf_s / wf_steps * 1e3,
                    "stages_ms": dict(zip(["sim", "negative_binomial_noise", "mean", "quantile_band"], stages)),
                    "h2d_bytes_per_step": int(draws_host.numel() * 8),
                    "d2h_bytes_per_step": int((mean_h.numel() + band_h.numel()) * 8),
                    "gpu_launches_per_step": 4 if world == 1 else 4 + 8,
                    "collectives_per_step": 0 if world == 1 else 7,
                    "what": "sim -> add_negative_binomial_noise(alpha=0.03) -> mean + 2.5 % / 97.5 % order statistics over the "
                            "draws of ALL ranks, on the device (sodm_noise, sodm_draws_sum, sodm_draws_order_stats / "
                            "sodm_draws_select_phase + NCCL all-reduce of sums and radix histograms)"}
        del ens, noisy

    # ---- max over ranks, final NCCL gather of per-rank summaries --------------------------------------------
    t = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        summary = torch.tensor([float(checksum)], dtype=torch.float64, device=dev)
        gathered = [torch.zeros_like(summary) for _ in range(world)]
        dist.all_gather(gathered, summary)          # the path's only collective: final gather (here of checksums)
    elapsed_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        value = world * n * args.steps / (elapsed_ms * 1e-3)
        e2e_value = world * n_e2e * e2e_steps / (e2e_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        k_ms = float(np.mean(kernel_ms))
        achieved = n * BYTES_PER_TRAJ / (k_ms * 1e-3) / 1e9
        prof = {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_inputs.json")))
        except OSError:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": prof.get("dram_bytes_per_launch"), "peak_source": peak_src,
                    "kernel": "sodm::tau_leap_kernel<InfluenzaJump,int,true,false>", "kernel_ms": k_ms,
                    "algorithmic_bytes_per_launch": n * BYTES_PER_TRAJ,
                    "note": "the kernel is ALU/issue-bound, not HBM-bound (7728 exact binomial variates per trajectory "
                            "vs 18 KB of output); see `alu`"}
        sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
        # issue-slot roofline: one warp instruction per cycle per SM sub-partition (148 SMs x 4) at the measured clock;
        # instructions per trajectory and the counters below come from the committed ncu capture of this kernel
        alu = {"binomial_variates_per_s": n * VARIATES_PER_TRAJ / (k_ms * 1e-3),
               "warp_inst_per_traj": prof.get("warp_inst_per_traj"),
               "issue_slot_peak_per_s": 148 * 4 * sm_clock * 1e6,
               "ncu_issue_active_pct": prof.get("issue_active_pct"), "ncu_lanes_per_inst": prof.get("lanes_per_inst"),
               "ncu_source": prof.get("source")}
        if prof.get("warp_inst_per_traj"):
            alu["achieved_warp_inst_per_s"] = prof["warp_inst_per_traj"] * n / (k_ms * 1e-3)
            alu["frac"] = alu["achieved_warp_inst_per_s"] / alu["issue_slot_peak_per_s"]
        roofline["alu"] = alu
        # CPU baseline beside it (N = 1 only): bounded sample on the host cores
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)        # the CPU arm gets every host core, not just the GPU's NUMA node
            os.environ.setdefault("OMP_NUM_THREADS", "1")
            rate, dt, cores = cpu_reference_rate(per_core=1024, rounds=1)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{cores} processes x 1024 trajectories of the same workload in {dt:.1f} s (oracle/jump.py, "
                             f"NumPy tau-leap vectorised over trajectories; the unmodified reference is ~5.6 traj/s/core)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "i32 counts / f32 rates and sampler (f64 exact test above mode 8192)",
                "data": "synthetic",
                "config": {"workload": "influenza_1718 tau-leap (BASELINE config 3): tau=0.5, 2017-12-01..2018-05-11 "
                                       "(322 steps, 162 outputs), 28 compartments, per-trajectory beta/f_ud draws",
                           "n_traj_per_gpu_per_step": n, "output": "int32 [28][162][n_traj] = 18.1 GB per step",
                           "l2_policy": "per-step output (18.1 GB) and inputs exceed the 126 MB L2; new seed each step",
                           "parallelism": f"trajectory-sharded x{world}, no data-path collective"},
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                        "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps, "n_traj_per_gpu_per_step": n_e2e,
                        "layout": "pinned host int32 (draws, time, age) per state, chunked double-buffered pipeline",
                        "chunk": args.chunk},
                "workflow": workflow,
                "gpu_launches": args.steps, "e2e_gpu_launches_per_step": e2e_launches // (e2e_warm + e2e_steps),
                "kernel_ms_per_step": kernel_ms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-traj", type=int, default=N_TRAJ)
    ap.add_argument("--e2e-traj", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workflow", action="store_true")
    ap.add_argument("--no-affinity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
