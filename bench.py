#!/usr/bin/env python
"""Headline benchmark: batched pendulum swing-up cubature i2c (BASELINE.json configs[2]:
4096 random initial states x T=200 per B200), metric = problem-timestep updates / s.

  python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the UNMODIFIED reference (oracle/_ref) on all host cores
  python bench.py --workload dcp|quadrotor|cartpole ...    # BASELINE configs[3] / configs[4] shards, cart-pole 4096 x 200 (fp64-bound rooflines)

One "step" = one EM iteration (I2cGraph.learn_msgs: forward + backward + M-step, i2c/i2c.py:1238-1245) over the
whole batch; one problem-timestep update = one cell through one such iteration.  Under torchrun every rank owns an
independent shard of problems (weak scaling, no data-path collective); the only collective is the final NCCL
gather of controllers / costs, timed in the end-to-end leg.
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
sys.path.insert(0, os.path.join(ROOT, "input-inference-for-control_b200"))

METRIC = "problem-timestep updates/sec (fp64, batched)"
UNIT = "updates/s"
# SURVEY.md section 8(d), pendulum row: algorithmic work per problem-timestep update
F_ALG, B_ALG = 3192.0, 704.0
# measured DRAM bytes per update of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture / updates in that launch): profiles/r01i_ncu_full_em_team_kernel_pendulum_4096.txt
# latency kernel (em_team_kernel<EnvPendulum,8,HOT>, 4096 x 200): profiles/r02l_ncu_full_em_team_kernel_pendulum_4096.txt
NCU_DRAM_BYTES_PER_UPDATE = (286.220544e6 + 399.501568e6) / (2 * 4096 * 200)
# throughput kernel: round 1 em_kernel<EnvPendulum,4> 509 B/update (profiles/r01d_ncu_full_em_kernel_pendulum_65536.txt); round 2:
NCU_DRAM_BYTES_PER_UPDATE_THROUGHPUT = 507.0  # em_ticket_kernel, profiles/r02_ncu_full_em_ticket_kernel_pendulum_65536.txt (16.61 GB / 32.768 M updates)
FP64_NOMINAL_TFLOPS = 37.2  # 148 SM x 64 FMA/clk x 2 x 1.965 GHz


def make_inputs(B, T, seed):
    """Config 3 of SURVEY.md 8(d): x0[b] = [pi,0] + [0.3,0.5] * N(0,I), mu_u[b] = 1e-2 N(0,1)."""
    rng = np.random.default_rng(seed)
    x0 = np.array([np.pi, 0.0]) + np.array([0.3, 0.5]) * rng.normal(size=(B, 2))
    mu_u = 1e-2 * rng.normal(size=(B, T, 1))
    return x0, mu_u


HYPER = dict(Q=np.diag([1.0, 100.0, 1.0]), R=np.diag([2.0]), alpha=100.0, tol=0.0, sig_u=2.0 * np.eye(1))


# ----------------------------------------------------------------------------- CPU arms
# kind "reference": the UNMODIFIED reference (oracle/_ref, a byte-for-byte copy made by oracle/install_ref.py; loaded through
# oracle/ref_shim.py) -- I2cGraph.learn_msgs, one problem after the other, one process per host core.
# kind "port": oracle/i2c_oracle.py, the batched NumPy restatement (about 60x faster per core than the reference itself).
def _port_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    from oracle import i2c_oracle as O

    x0, mu_u, steps, warmup = args
    g = O.make_graph("PendulumKnown", mu_u.shape[1], HYPER["Q"], HYPER["R"], HYPER["Q"], HYPER["alpha"], HYPER["tol"], mu_u,
                     HYPER["sig_u"], B=x0.shape[0], x0=x0)
    for _ in range(warmup):
        g.learn_msgs()
    t0 = time.perf_counter()
    for _ in range(steps):
        g.learn_msgs()
    return time.perf_counter() - t0, x0.shape[0], 0.0


def cpu_rate(kind, B, T, steps, warmup, per_core, cores, seed=1234):
    """`cores` processes, `per_core` problems each, taken (evenly strided) from the SAME seeded inputs as the CUDA arm's
    rank 0.  Returns (updates/s, slowest worker's seconds, wall seconds, sample description)."""
    import multiprocessing as mp

    x0, mu_u = make_inputs(B, T, seed)
    idx = np.linspace(0, B - 1, cores * per_core).astype(int)
    if kind == "reference":
        from oracle import ref_bench

        fn = ref_bench.em_worker
        jobs = [(x0[idx[c::cores]], mu_u[idx[c::cores]], HYPER, steps, warmup) for c in range(cores)]
    else:
        fn = _port_worker
        jobs = [(x0[idx[c::cores]], mu_u[idx[c::cores]], steps, warmup) for c in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        out = pool.map(fn, jobs)
        wall = time.perf_counter() - t0
    dt = max(o[0] for o in out)
    n = sum(o[1] for o in out)
    what = ("unmodified reference I2cGraph.learn_msgs (oracle/_ref via oracle/ref_shim.py)" if kind == "reference"
            else "oracle/i2c_oracle.py, batched NumPy restatement")
    sample = (f"{cores} procs x {per_core} problems (evenly strided from the {B}-problem inputs, seed {seed}) x T={T}, "
              f"{steps} EM iterations after {warmup} warm-up; {what}")
    return n * T * steps / dt, dt, wall, sample


def cpu_kind():
    try:
        from oracle import ref_bench

        return "reference" if ref_bench.available() else "port"
    except Exception:
        return "port"


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    T, K, W = args.horizon, args.steps, min(args.warmup, 1)
    # bounded sample: about one minute per worker (the reference does ~1.2e3 updates/s/core, the port ~7e4)
    est = 1.2e3 if kind == "reference" else 7e4
    per_core = int(max(1, min(64, 60.0 * est / (T * (K + W)))))
    rate, dt, wall, sample = cpu_rate(kind, args.problems, T, K, W, per_core, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"pendulum swing-up cubature i2c, {args.problems} problems/GPU x T={T} (BASELINE configs[2])",
                   "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "seconds": wall},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


# ----------------------------------------------------------------------------- CUDA arm
def bind_to_gpu_numa_node(dev):
    """One process per GPU: run (and allocate pinned host buffers, first touch) on the CPUs NVML reports as local to this
    GPU, so that the H2D / D2H copies of eight ranks do not cross the socket interconnect."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_cuda(args, rank, world, local_rank):
    import torch
    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist.barrier()
    if rank != 0:
        ge.build()
    import i2c_b200
    from i2c_b200 import capi

    dev = local_rank
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa_node(dev) if world > 1 else None
    B, T, K, W = args.problems, args.horizon, args.steps, args.warmup
    x0, mu_u = make_inputs(B, T, 1234 + rank)
    g = i2c_b200.BatchedI2c("PendulumKnown", B, T, HYPER["Q"], HYPER["R"], HYPER["Q"], HYPER["alpha"], HYPER["tol"], mu_u,
                            HYPER["sig_u"], x0=x0, device=dev, max_iters=max(K, W, 1))

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput: K EM iterations, inputs already in HBM, one persistent launch
    g.run(max(W, 3), capi.PH_LEARN, collect=False)  # warm-up (>= 3 iterations)
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = g.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.perf_counter()
    e0.record()
    g.run(K, capi.PH_LEARN, collect=False)
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = g.kernel_launches() - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    kernel_ms = max_over_ranks(g.last_run_ms())
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    st, _ = g.status()
    n_fail = int(np.count_nonzero(st))
    value = world * B * T * K / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers.  Every step uploads the start-state belief (sys.x0 /
    # sys.sig_x0 are re-read by every sweep of the reference, i2c.py:876-880), runs one learn_msgs and reads the step's
    # result -- cost and alpha of every problem -- back to the host.  The controllers stay in HBM between the EM iterations
    # and cross once, after the last step: each rank reads its own K, k, sigK into pinned memory and (N > 1) the ranks
    # all_gather controllers and costs over NVLink -- the design BASELINE.json's north star describes ("NCCL ... solely for
    # the final gather of controllers and costs").  All of it is inside the timed region.
    Ke = args.e2e_steps if args.e2e_steps > 0 else max(3, K)  # default 100 = the EM iterations of BASELINE configs[2]'s job
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()  # noqa: E731
    h_x0, h_s0 = pin(B, 2), pin(B, 2, 2)
    h_x0[:], h_s0[:] = g.x0, g.sig_x0
    h_m2 = [pin(2, B), pin(2, B)]
    h_m = h_m2[0]
    h_pol = [(pin(B, T, 1, 2), pin(B, T, 1), pin(B, T, 1, 1)) for _ in range(2)]
    h_K, h_k, h_s = h_pol[0]
    L = g.lib

    m_ids = np.array([capi.METRICS["alpha"], capi.METRICS["cost_m"]], np.int32)

    def step_metrics_only():
        capi.check(L.i2c_set_initial_state_async(g._h, capi.ptr(h_x0), capi.ptr(h_s0)))  # pinned, persistent buffers
        capi.check(L.i2c_run(g._h, 1, capi.PH_LEARN))
        capi.check(L.i2c_get_metrics(g._h, capi.ptr(m_ids), 2, capi.ptr(h_m), 1))  # alpha -> h_m[0], cost -> h_m[1]; one sync

    log = np.zeros((2, 2))  # what the host does with a step's result (scripts/i2c_run.py:84-88 prints cost and alpha)

    def step_pipelined(i):
        # step i is queued (belief upload, learn_msgs, read-back of cost + alpha through staging slot i & 1) BEFORE the host
        # collects step i-1: the stream never drains, the D2H of a step overlaps the next step's kernel
        capi.check(L.i2c_set_initial_state_async(g._h, capi.ptr(h_x0), capi.ptr(h_s0)))
        capi.check(L.i2c_run(g._h, 1, capi.PH_LEARN))
        capi.check(L.i2c_get_last_metrics_async(g._h, capi.ptr(m_ids), 2, capi.ptr(h_m2[i & 1]), i & 1))
        if i > 0:
            capi.check(L.i2c_metrics_wait(g._h, (i - 1) & 1))
            log[(i - 1) & 1] = h_m2[(i - 1) & 1][:, 0]

    def final_gather():
        # the path's only collective: final gather of controllers and costs over NVLink
        from i2c_b200 import dist as idist

        Kd, kd, sd = g.policy_device_tensors()
        cost = torch.from_numpy(np.ascontiguousarray(h_m2[(Ke - 1) & 1][1])).to(Kd.device)
        gathered = idist.gather_controllers(Kd, kd, sd, world * B, extra=(cost,), dst=0)  # onto rank 0 (it writes the results)
        assert (gathered[0].shape[0] == world * B) if rank == 0 else (gathered[0] is None)
        torch.cuda.synchronize(dev)

    parts = {}

    def e2e_run(n):
        ta = time.perf_counter()
        for i in range(n):
            step_pipelined(i)
        capi.check(L.i2c_metrics_wait(g._h, (n - 1) & 1))
        log[(n - 1) & 1] = h_m2[(n - 1) & 1][:, 0]
        tb = time.perf_counter()
        # the rank's own controllers go to pinned host memory on the copy stream (PCIe) WHILE the ranks gather over NVLink
        capi.check(L.i2c_get_policy_async(g._h, capi.ptr(h_K), capi.ptr(h_k), capi.ptr(h_s)))
        if dist is not None:
            final_gather()
        tc = time.perf_counter()
        capi.check(L.i2c_copy_wait(g._h))
        td = time.perf_counter()
        parts.update(loop_ms=(tb - ta) * 1e3, gather_ms=(tc - tb) * 1e3, final_d2h_wait_ms=(td - tc) * 1e3)

    e2e_run(2)  # warm-up (communicator buffers, allocator, pinned pages)
    barrier()
    t0 = time.perf_counter()
    e2e_run(Ke)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    parts = {k: max_over_ranks(v) for k, v in sorted(parts.items())}
    e2e_value = world * B * T * Ke / (e2e_ms * 1e-3)
    h2d = h_x0.nbytes + h_s0.nbytes
    d2h_step = h_m.nbytes
    d2h_once = h_K.nbytes + h_k.nbytes + h_s.nbytes

    # the same loop with a host synchronisation in every step (i2c_get_metrics: the host sees cost / alpha before it queues the
    # next step), side key
    Ks = max(3, min(Ke, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ks):
        step_metrics_only()
    barrier()
    e2es_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2es_value = world * B * T * Ks / (e2es_ms * 1e-3)

    # The reference's scripts/i2c_run.py additionally reads the controller after EVERY iteration (:89-98, for its roll-out
    # evaluation on the host); the same loop with that per-step 26 MB read-back, for comparison (side key):
    def step_with_policy(i):
        step_metrics_only()
        bK, bk, bs = h_pol[i & 1]
        capi.check(L.i2c_get_policy_async(g._h, capi.ptr(bK), capi.ptr(bk), capi.ptr(bs)))

    Kp = max(3, min(K, 10))
    step_with_policy(0)
    capi.check(L.i2c_copy_wait(g._h))
    barrier()
    t0 = time.perf_counter()
    for i in range(Kp):
        step_with_policy(i)
    capi.check(L.i2c_copy_wait(g._h))
    barrier()
    e2ep_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2ep_value = world * B * T * Kp / (e2ep_ms * 1e-3)

    mpc = mpc_leg(args, dev, rank, world, max_over_ranks, barrier) if args.mpc_rollouts > 0 else None
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    per_gpu_rate = B * T * K / (kernel_ms * 1e-3)
    achieved = B_ALG * per_gpu_rate / 1e9
    try:
        fp64_peak = capi.dfma_peak(dev)
    except Exception:
        fp64_peak = None
    fp64_ach = F_ALG * per_gpu_rate / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"pendulum swing-up cubature i2c, {B} problems/GPU x T={T} (BASELINE configs[2])",
                   "problems_per_gpu": B, "horizon": T, "em_iterations_timed": K, "updates_per_step": world * B * T,
                   "l2": "per-iteration working set (prior+posterior+filtered records) = %.0f MB > 126 MB L2; no flush needed"
                         % ((13 + 13 + 20) * 8 * B * T / 1e6),
                   "launch": "one persistent kernel launch runs all K iterations (forward, backward, M-step fused)",
                   "failed_problems": n_fail},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_step,
                "d2h_bytes_once": d2h_once, "steps": Ke, "ms_per_step": e2e_ms / Ke, "breakdown_ms_max_over_ranks": parts,
                "what": "per step: H2D start-state belief (pinned), one learn_msgs through the C-ABI, D2H cost + alpha of every "
                        "problem (pipelined: step i+1 is queued before the host collects step i, i2c_get_last_metrics_async); after the last step, inside the timed region: D2H of K, k, sigK of this rank "
                        "(once, on the copy stream)" + ("; concurrently the NCCL gather of controllers and costs onto rank 0 over NVLink" if world > 1 else "")
                        + (f"; process bound to the {numa} CPUs local to its GPU" if numa else "")},
        "e2e_synchronous_steps": {"value": e2es_value, "unit": UNIT, "steps": Ks, "ms_per_step": e2es_ms / Ks,
                                  "what": "as e2e without the final controller read-back, but the host waits for cost + alpha of "
                                          "step i before it queues step i+1 (i2c_get_metrics): one stream drain per step"},
        "e2e_policy_every_step": {"value": e2ep_value, "unit": UNIT, "steps": Kp, "h2d_bytes_per_step": h2d,
                                  "d2h_bytes_per_step": d2h_step + d2h_once,
                                  "what": "as e2e, but K, k, sigK are read back after EVERY step (scripts/i2c_run.py:89-98 does "
                                          "that for its host-side roll-out evaluation): bound by 26 MB/step over PCIe"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": NCU_DRAM_BYTES_PER_UPDATE * B * T * K if (B == 4096 and T == 200) else None,
                     "traffic_source": "ncu dram bytes/update (profiles/r02l_ncu_full_em_team_kernel_pendulum_4096.txt; final build: r02z, 418 B) x updates "
                                       "per launch; algorithmic bytes per launch = %.4g" % (B_ALG * B * T * K),
                     "peak_source": peak_src,
                     "kernel": "em_team_kernel<EnvPendulum,8,HOT>" if B <= 148 * 32 else "em_kernel<EnvPendulum>",
                     "kernel_ms_per_launch": kernel_ms,
                     "algorithmic_bytes_per_update": B_ALG, "algorithmic_flops_per_update": F_ALG,
                     "fp64": {"achieved_tflops": fp64_ach, "peak_measured_tflops": fp64_peak,
                              "frac_of_measured": (fp64_ach / fp64_peak) if fp64_peak else None,
                              "peak_nominal_tflops": FP64_NOMINAL_TFLOPS, "frac_of_nominal": fp64_ach / FP64_NOMINAL_TFLOPS}},
    }
    if world == 1 and args.saturation > B:
        # same path at batches that fill the machine (the 4096-problem headline is bound by the latency of the sequential
        # recursion over t: 128 warps on 592 sub-partitions): the B at which >= 50 % of the roofline is reached, and the
        # saturated rate.  Kernel by batch size: <= 4736 team kernel (8 warps per tile), <= 9472 team kernel (4 warps),
        # < 56832 one warp per tile (255 registers), >= 56832 em_ticket_kernel ((tile, iteration) work items, no wave edges)
        del g
        torch.cuda.empty_cache()

        def at_batch(Bs, n):
            x0s, mu_us = make_inputs(Bs, T, 99)
            gs = i2c_b200.BatchedI2c("PendulumKnown", Bs, T, HYPER["Q"], HYPER["R"], HYPER["Q"], HYPER["alpha"], HYPER["tol"],
                                     mu_us, HYPER["sig_u"], x0=x0s, device=dev, max_iters=max(8, n))
            gs.run(3, capi.PH_LEARN, collect=False)
            gs.run(n, capi.PH_LEARN, collect=False)
            torch.cuda.synchronize(dev)
            ms_s = gs.last_run_ms()
            failed = int(np.count_nonzero(gs.status()[0]))
            gs.close()
            del gs
            torch.cuda.empty_cache()
            return ms_s / n, Bs * T * n / (ms_s * 1e-3), failed

        sweep = []
        for Bs in args.b_sweep:
            ms_i, rate_i, failed = at_batch(Bs, 10)
            sweep.append({"problems": Bs, "ms_per_step": ms_i, "value": rate_i, "hbm_frac": B_ALG * rate_i / 1e9 / hbm_peak,
                          "failed_problems": failed})
        line["b_sweep"] = {"unit": UNIT, "horizon": T, "em_iterations_timed": 10, "hbm_frac": "algorithmic bytes (704 B/update) x "
                           "rate / measured HBM peak", "points": sweep}
        Bs = args.saturation
        ms_i, rate_s, failed = at_batch(Bs, 10)
        line["saturation"] = {"problems": Bs, "value": rate_s, "unit": UNIT, "ms_per_step": ms_i,
                              "hbm_frac_by_algorithmic_bytes": B_ALG * rate_s / 1e9 / hbm_peak,
                              "hbm_frac_by_measured_traffic": NCU_DRAM_BYTES_PER_UPDATE_THROUGHPUT * rate_s / 1e9 / hbm_peak,
                              "hbm_frac_note": "by ALGORITHMIC bytes (704 B/update); em_ticket_kernel moves 507 B/update "
                                               "(ncu r02, packed triangles), i.e. %.2f of the measured copy bandwidth"
                                               % (NCU_DRAM_BYTES_PER_UPDATE_THROUGHPUT * rate_s / 1e9 / hbm_peak),
                              "fp64_frac_of_measured": (F_ALG * rate_s / 1e12 / fp64_peak) if fp64_peak else None,
                              "failed_problems": failed}
    if mpc is not None:
        line["mpc"] = mpc
    line["em_iterations_per_s"] = {"batch_sweeps_per_s": 1e3 * K / ms, "problem_iterations_per_s": world * B * K / (ms * 1e-3)}
    if world == 1 and args.scan_horizon > 0:
        line["time_parallel"] = scan_leg(args, dev)
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        kind = cpu_kind()
        if kind == "reference":
            # ~15-25 s per core: 8 problems x (1 + 10) iterations x T at 1-2e3 updates/s/core
            rate, dt, wall, sample = cpu_rate("reference", B, T, 10, 1, 8, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                                    "seconds": wall}
        prate, pdt, pwall, psample = cpu_rate("port", B, T, 20, 1, 64, cores)
        port = {"value": prate, "unit": UNIT, "cores": cores, "kind": "port", "sample": psample, "seconds": pwall}
        if kind == "reference":
            line["cpu_baseline_port"] = port
        else:
            line["cpu_baseline"] = port
        if mpc is not None:
            line["mpc"]["cpu_baseline"] = mpc_cpu_baseline(cores)
        try:
            line["config1"] = config1_leg()
        except Exception as e:  # a side leg must never cost the headline line
            line["config1"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def config1_leg(n_iter=200, n_ref=5):
    """BASELINE configs[0]: scripts/i2c_run.py pendulum_known_quad -- ONE problem, T = 100, the script's own loop
    (i2c_run.py:84-98: learn_msgs, then both controller getters, every iteration) through the drop-in mirror
    (`from i2c.i2c import I2cGraph`, B = 1: one tile, one block) against the unmodified reference on one host core.
    A single problem cannot fill a GPU: this is the latency of the drop-in path, not a throughput number."""
    import sys as _sys

    pkg = os.path.join(ROOT, "input-inference-for-control_b200")
    if pkg not in _sys.path:
        _sys.path.insert(0, pkg)
    from i2c.exp_types import CubatureQuadrature
    from i2c.i2c import I2cGraph
    from i2c.model import make_env_model
    from i2c.policy.linear import ExpertTimeIndexedLinearGaussianPolicy, TimeIndexedLinearGaussianPolicy

    T = 100
    Q, R, Qf = np.diag([1.0, 100.0, 1.0]), np.diag([2.0]), np.diag([1.0, 100.0, 1.0])  # experiments/pendulum_known_quad.py
    rng = np.random.default_rng(0)
    mu_u = 1e-2 * rng.normal(size=(T, 1))
    model = make_env_model("PendulumKnown", None)
    g = I2cGraph(model, T, Q, R, Qf, 100.0, 0.0, mu_u, 2.0 * np.eye(1), None, None, CubatureQuadrature(1, 0, 0), res_dir=None)
    pl = TimeIndexedLinearGaussianPolicy(0.0 * np.eye(1), T, 1, 2)
    pe = ExpertTimeIndexedLinearGaussianPolicy(0.0 * np.eye(1), T, 1, 2, soft=False)
    g.reset_metrics()

    def it():
        g.learn_msgs()
        pl.write(*g.get_local_linear_policy())
        pe.write(*g.get_local_expert_linear_policy())

    for _ in range(3):
        it()
    t0 = time.perf_counter()
    for _ in range(n_iter):
        it()
    ms = (time.perf_counter() - t0) * 1e3 / n_iter
    t0 = time.perf_counter()
    for _ in range(n_iter):
        g.learn_msgs()
    ms_learn = (time.perf_counter() - t0) * 1e3 / n_iter
    out = {"workload": "scripts/i2c_run.py pendulum_known_quad: 1 problem x T=100, learn_msgs + both controller getters per "
                       "iteration through the drop-in mirror (BASELINE configs[0])",
           "ms_per_em_iteration": ms, "ms_per_learn_msgs_only": ms_learn, "em_iterations_timed": n_iter,
           "final_alpha": float(g.alphas[-1])}
    if cpu_kind() == "reference":
        from oracle import ref_shim

        ns = ref_shim.load()
        sys_ = ns.model.make_env_model("PendulumKnown", None)
        r = ns.i2c.I2cGraph(sys_, T, Q, R, Qf, 100.0, 0.0, mu_u, 2.0 * np.eye(1), None, None, ns.exp_types.CubatureQuadrature(1, 0, 0))
        r.learn_msgs()
        t0 = time.perf_counter()
        for _ in range(n_ref):
            r.learn_msgs()
            r.get_local_linear_policy()
            r.get_local_expert_linear_policy()
        out["cpu_baseline"] = {"value": (time.perf_counter() - t0) * 1e3 / n_ref, "unit": "ms per EM iteration", "cores": 1,
                               "kind": "reference", "sample": f"{n_ref} iterations of the unmodified reference after 1 warm-up"}
        out["speedup_vs_reference"] = out["cpu_baseline"]["value"] / ms
        ref_shim.unload()
    return out


def scan_leg(args, dev):
    """Parallel-in-time variant (i2c_run_scan, csrc/i2c_scan.cuh) on a long-horizon linear-Gaussian problem: the horizon
    is cut into chunks that are filtered / smoothed concurrently (associative Kalman / RTS elements).  Sequential kernel
    beside it; both produce the same records (tests/test_gpu_scan.py)."""
    import i2c_b200

    B, T, chunk = 32, args.scan_horizon, 64
    rng = np.random.default_rng(5)
    A = np.array([[1.0, 0.1], [-0.05, 0.98]]) + 0.01 * rng.normal(size=(B, 2, 2))
    xg = rng.normal(size=(B, 2))
    par = i2c_b200.envs.linear_params(A, np.array([[0.0], [0.1]]), xg - np.einsum("bij,bj->bi", A, xg))
    z = np.repeat(np.concatenate((xg, np.zeros((B, 1))), axis=1)[:, None, :], T, axis=1)
    out = {"problems": B, "horizon": T, "chunk_cells": chunk, "inference": "Linearize, LinearKnown (exact scan)"}
    ref = None
    for name, ch in (("sequential", None), ("scan", chunk)):
        G = i2c_b200.BatchedI2c("LinearKnown", B, T, np.diag([1.0, 2.0]), np.diag([0.5]), np.diag([1.0, 2.0]), 5.0, 0.5,
                                1e-2 * rng.normal(size=(B, T, 1)) * 0, np.eye(1), x0=xg + 2.0, sig_x0=1e-2 * np.eye(2),
                                sig_eta=1e-3 * np.eye(2), env_par=par, z=z, z_term=xg, z_per_problem=True, inference="linearize",
                                device=dev, max_iters=8)
        G.time_parallel_chunk = ch
        G.forward_backward(2)
        G.synchronize()
        G.forward_backward(5)
        G.synchronize()
        out[f"{name}_ms_per_sweep_pair"] = G.last_run_ms() / 5
        K = G.field("K")
        if ref is None:
            ref = K
        else:
            out["max_rel_diff_K"] = float(np.max(np.abs(K - ref)) / np.max(np.abs(ref)))
        G.close()
    out["speedup"] = out["sequential_ms_per_sweep_pair"] / out["scan_ms_per_sweep_pair"]
    return out


def mpc_leg(args, dev, rank, world, max_over_ranks, barrier, n_timed=20, with_em_timing=False):
    """Second half of BASELINE.json's metric: microseconds per MPC solve.  Quadrotor MPC with cubature Kalman filter
    state estimation (BASELINE configs[4], scripts/mpc_state_est/mpc_quad.py:538-652): `mpc_rollouts` parallel
    closed-loop roll-outs per GPU, T_plan = 10, mpc_iter = 2; one solve = PartiallyObservedMpcPolicy.__call__ =
    CKF step + 2 x (forward + backward sweep, _update_priors) + first action + horizon shift, through the public
    API with HOST measurement / action buffers (so H2D / D2H are inside the timed region)."""
    import torch
    import i2c_b200

    B = args.mpc_rollouts
    W_, H_ = i2c_b200.envs.QUAD_W, i2c_b200.envs.QUAD_H
    T, T_plan, mpc_iter = 100, 10, 2
    z_traj = np.zeros((T, 8))
    z_traj[:, 0] = np.linspace(W_ / 4, 3 * W_ / 4, T)
    z_traj[:, 1] = H_ / 2 + (H_ / 4) * np.sin(np.linspace(0, 2 * np.pi, T))
    z_traj[:, 2] = 2 * np.pi * np.heaviside(np.linspace(-1, 1, T), 1)
    Q, R = np.diag([1e3, 1e3, 1e3, 1, 1, 1]), np.diag([1e-3, 1e-3])
    sig_zeta = np.diag([1e-6] * 8)
    u_init = 0.5 * 9.81 * i2c_b200.envs.QUAD_MASS * np.ones((T_plan, 2))
    g = i2c_b200.BatchedI2c("Quadrotor", B, T_plan, Q, R, Q / 1e3, 1.0, 1.0, u_init, 1e-2 * np.eye(2), device=dev)
    g._propagate = True
    pol = i2c_b200.BatchedPartiallyObservedMpc(g, mpc_iter, 1e-2 * np.eye(2), z_traj, sig_zeta=sig_zeta, pinned_io=True)
    pol.set_control(feedforward=False)
    g.calibrate_alpha()
    pol.optimize(25)
    g.calibrate_alpha()
    rng = np.random.default_rng(7 + rank)
    e = g.env
    y0 = np.array([e.x0[0] - 0.8, e.x0[1], e.x0[0] + 0.8, e.x0[1], 0, 0, 0.8, 0.8])  # measure(x0)
    n_warm = 3
    # synthetic measurement stream, generated before the timed region (host RNG is not part of the solve)
    ys = torch.empty((n_warm + n_timed, B, 8), dtype=torch.float64, pin_memory=True).numpy()
    ys[:] = y0 + 1e-3 * rng.normal(size=ys.shape)
    u = np.zeros((B, 2))
    launches0 = 0
    t0 = 0.0
    for t in range(n_warm + n_timed):
        if t == n_warm:
            barrier()
            launches0 = g.kernel_launches()
            t0 = time.perf_counter()
        u = pol(t, ys[t], u)  # one fused library call: CKF + 2 sweeps + first action + horizon shift (H2D y,u; D2H u)
    barrier()
    ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_timed
    st = g.status()[0]
    em_ms = None
    if with_em_timing:
        # the EM sweeps of one solve on their own (CUDA events around the kernel): 2 x (forward + backward, _update_priors)
        from i2c_b200 import capi

        g.run(mpc_iter, capi.PH_FORWARD | capi.PH_BACKWARD | capi.PH_UPDATE_PRIORS, collect=False)
        g.run(mpc_iter, capi.PH_FORWARD | capi.PH_BACKWARD | capi.PH_UPDATE_PRIORS, collect=False)
        g.synchronize()
        em_ms = max_over_ranks(g.last_run_ms())
    return {"em_kernel_ms_per_control_step": em_ms, "metric": "us per MPC solve (amortised over the roll-out batch)", "value": ms * 1e3 / (world * B), "unit": "us",
            "higher_is_better": False, "rollouts_per_gpu": B, "batch_latency_ms_per_control_step": ms,
            "solves_per_s": world * B / (ms * 1e-3), "control_steps_timed": n_timed,
            "launches_per_control_step": (g.kernel_launches() - launches0) / n_timed,
            "config": "quadrotor MPC + cubature Kalman filter, T_plan=10, mpc_iter=2, feedback mode (BASELINE configs[4])",
            "failed_rollouts": int(np.count_nonzero(st))}


# ----------------------------------------------------------------------------- BASELINE configs[3] / configs[4]
ALG = {"DoubleCartpoleKnown": (34473.0, 3400.0), "Quadrotor": (33105.0, 4160.0), "CartpoleKnown": (11601.0, 1792.0)}  # (flops, bytes) per update, SURVEY.md 8(d)


def _peaks(capi, dev):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    try:
        fp64 = capi.dfma_peak(dev)
        src = "measured on this GPU (i2c_dfma_peak: 8 independent DFMA chains per thread)"
    except Exception:
        fp64, src = FP64_NOMINAL_TFLOPS, "nominal 37.2 TFLOP/s"
    return hbm, fp64, src


def run_workload(args, rank, world, local_rank):
    """`--workload dcp`: double cart-pole cubature i2c + covariance control, 2048 problems/GPU x T=500 (= 16384 over 8 GPUs,
    BASELINE configs[3]).  The reference algorithm cannot run the in-loop propagate on these inputs (DESIGN.md section 4:
    the propagation of the first posterior loses positive definiteness in the reference itself), so the timed EM iteration
    is forward + backward + M-step with covariance control, WITHOUT the in-loop propagate -- stated in config.workload.
    `--workload quadrotor`: quadrotor MPC with cubature Kalman filter, 8192 roll-outs/GPU (= 65536 over 8, configs[4]);
    the line's metric is microseconds per MPC solve; the roofline is that of the EM sweeps inside the solve.
    Both are fp64-pipe-bound: the roofline denominator is the measured DFMA peak."""
    import torch
    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dist.barrier()
    if rank != 0:
        ge.build()
    import i2c_b200
    from i2c_b200 import capi

    dev = local_rank
    torch.cuda.set_device(dev)
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(dev)
    if args.workload == "quadrotor":
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        t_start = time.perf_counter()
        mpc = mpc_leg(args, dev, rank, world, max_over_ranks, barrier, n_timed=max(K, 10), with_em_timing=True)
        t_end = time.perf_counter()
        clocks = sampler.stop(t_start, t_end) if rank == 0 else None
        if rank != 0:
            return
        hbm, fp64, src = _peaks(capi, dev)
        F, Bb = ALG["Quadrotor"]
        upd = mpc["rollouts_per_gpu"] * 10 * 2  # cells x mpc_iter sweeps per solve batch
        rate = upd / (mpc["em_kernel_ms_per_control_step"] * 1e-3)
        line = {"metric": "us per MPC solve (fp64, batched; amortised over the closed-loop roll-outs)", "value": mpc["value"],
                "unit": "us", "n_gpus": world, "steps": mpc["control_steps_timed"], "warmup": 3,
                "ms_per_step": mpc["batch_latency_ms_per_control_step"], "higher_is_better": False, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"quadrotor MPC + cubature Kalman filter (mpc_quad.py), {mpc['rollouts_per_gpu']} closed-loop "
                                       f"roll-outs/GPU, T_plan=10, mpc_iter=2, feedback mode (BASELINE configs[4])",
                           "l2": "records of 8192 roll-outs x 10 cells = 240 MB > 126 MB L2", "failed": mpc["failed_rollouts"]},
                "clocks": clocks,
                "e2e": {"value": mpc["value"], "unit": "us", "h2d_bytes_per_step": mpc["rollouts_per_gpu"] * (8 + 2) * 8,
                        "d2h_bytes_per_step": mpc["rollouts_per_gpu"] * 2 * 8,
                        "what": "the line's value IS end to end: every control step uploads measurements + applied actions from "
                                "pinned host memory and reads the new actions back (one fused i2c_mpc_step call)"},
                "gpu_launches": int(mpc["launches_per_control_step"] * mpc["control_steps_timed"]),
                "roofline": {"bound": "fp64", "achieved": F * rate / 1e12, "peak": fp64, "unit": "TFLOP/s",
                             "frac": F * rate / 1e12 / fp64, "traffic": None, "peak_source": src,
                             "kernel": "em_team_kernel / em_kernel<EnvQuadrotor> (2 sweeps x 10 cells per solve)",
                             "kernel_ms_per_launch": mpc["em_kernel_ms_per_control_step"],
                             "algorithmic_flops_per_update": F, "algorithmic_bytes_per_update": Bb,
                             "hbm_frac": Bb * rate / 1e9 / hbm},
                "mpc": mpc}
        if world == 1 and not args.no_cpu_baseline:
            cb = mpc_cpu_baseline(os.cpu_count() or 1)
            if cb:
                line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return

    rng = np.random.default_rng(4321 + rank)
    if args.workload == "cartpole":
        # ---- cart-pole swing-up (the north star's second target system; hyper-parameters of cartpole_known_quad.py:23-34)
        env = "CartpoleKnown"
        B, T = args.problems, args.horizon
        e = i2c_b200.envs.make(env)
        Q, R = np.diag([1.0, 1.0, 100.0, 10.0, 1.0]), np.diag([1.0])
        x0 = e.x0 + 0.05 * rng.normal(size=(B, 4))
        mu_u = 1e-2 * rng.normal(size=(B, T, 1))
        g = i2c_b200.BatchedI2c(env, B, T, Q, R, Q, 80.0, 0.0, mu_u, np.eye(1), x0=x0, device=dev, max_iters=max(K, W, 1))
        what = (f"cart-pole swing-up cubature i2c, {B} problems/GPU x T={T} (the north star's second target system; "
                f"hyper-parameters of cartpole_known_quad.py)")
        n_rec = 26 + 26 + 54
    else:
        # ---- double cart-pole, covariance control
        env = "DoubleCartpoleKnown"
        B, T = args.problems if args.problems != 4096 else 2048, args.horizon if args.horizon != 200 else 500
        sf = 1e-3
        Q = sf * np.diag([1.0, 1.0, 100.0, 1.0, 100.0, 10.0, 1.0, 1.0])
        R = sf * np.diag([0.1])
        mu_t, sig_t = np.zeros(6), np.diag([0.01, 0.005, 0.005, 0.05, 0.05, 0.05])
        e = i2c_b200.envs.make(env)
        x0 = e.x0 + 0.05 * rng.normal(size=(B, 6))
        mu_u = 1e-2 * rng.normal(size=(B, T, 1))
        g = i2c_b200.BatchedI2c(env, B, T, Q, R, Q, 0.05, 0.99, mu_u, np.eye(1), mu_t, sig_t, x0=x0, device=dev,
                                max_iters=max(K, W, 1))
        g.set_cell_flag(capi.CELL_EXPERT, False)
        g.propagate()  # the initial propagate of nonlinear_covariance_control.py:105-113
        what = (f"double cart-pole cubature i2c + covariance control, {B} problems/GPU x T={T} (BASELINE "
                f"configs[3] = 16384 x 500 over 8 GPUs); EM iteration = forward + backward + M-step WITHOUT the "
                f"in-loop propagate (infeasible for the reference algorithm on these inputs, DESIGN.md section 4); "
                f"initial propagate() run once before")
        n_rec = 44 + 44 + 104
    dx = e.dim_x
    g.run(W, capi.PH_LEARN, collect=False)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = g.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.perf_counter()
    e0.record()
    g.run(K, capi.PH_LEARN, collect=False)
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = g.kernel_launches() - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    kernel_ms = max_over_ranks(g.last_run_ms())
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    n_fail = int(np.count_nonzero(g.status()[0]))
    # end to end: per step H2D belief + one learn_msgs + D2H cost / alpha; controllers once at the end
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()  # noqa: E731
    h_x0, h_s0, h_m = pin(B, dx), pin(B, dx, dx), pin(2, B)
    h_x0[:], h_s0[:] = g.x0, g.sig_x0
    h_K, h_k, h_s = pin(B, T, 1, dx), pin(B, T, 1), pin(B, T, 1, 1)
    L = g.lib
    Ke = max(3, min(K, 10))

    m_ids = np.array([capi.METRICS["alpha"], capi.METRICS["cost_m"]], np.int32)
    h_m2 = [h_m, pin(2, B)]

    def e2e_run(n):
        # as the pendulum line: belief upload into the alternate buffers, one learn_msgs, pipelined read of cost + alpha
        for i in range(n):
            capi.check(L.i2c_set_initial_state_async(g._h, capi.ptr(h_x0), capi.ptr(h_s0)))
            capi.check(L.i2c_run(g._h, 1, capi.PH_LEARN))
            capi.check(L.i2c_get_last_metrics_async(g._h, capi.ptr(m_ids), 2, capi.ptr(h_m2[i & 1]), i & 1))
            if i > 0:
                capi.check(L.i2c_metrics_wait(g._h, (i - 1) & 1))
        capi.check(L.i2c_metrics_wait(g._h, (n - 1) & 1))
        capi.check(L.i2c_get_policy_async(g._h, capi.ptr(h_K), capi.ptr(h_k), capi.ptr(h_s)))
        if dist is not None:
            from i2c_b200 import dist as idist

            Kd, kd, sd = g.policy_device_tensors()
            idist.gather_controllers(Kd, kd, sd, world * B, dst=0)
            torch.cuda.synchronize(dev)
        capi.check(L.i2c_copy_wait(g._h))

    e2e_run(1)
    barrier()
    t0 = time.perf_counter()
    e2e_run(Ke)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    if rank != 0:
        return
    hbm, fp64, src = _peaks(capi, dev)
    F, Bb = ALG[env]
    rate_gpu = B * T * K / (kernel_ms * 1e-3)
    line = {"metric": METRIC, "value": world * B * T * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": what,
                       "l2": "records of one iteration = %.0f MB > 126 MB L2" % (n_rec * 8 * B * T / 1e6),
                       "failed_problems": n_fail},
            "clocks": clocks,
            "e2e": {"value": world * B * T * Ke / (e2e_ms * 1e-3), "unit": UNIT, "steps": Ke,
                    "h2d_bytes_per_step": h_x0.nbytes + h_s0.nbytes, "d2h_bytes_per_step": h_m.nbytes,
                    "d2h_bytes_once": h_K.nbytes + h_k.nbytes + h_s.nbytes,
                    "what": "per step H2D belief + learn_msgs + D2H cost / alpha (pipelined); K, k, sigK read back once after the last "
                            "step" + (", overlapped with the NCCL gather onto rank 0" if world > 1 else "")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": F * rate_gpu / 1e12, "peak": fp64, "unit": "TFLOP/s",
                         "frac": F * rate_gpu / 1e12 / fp64, "traffic": None, "peak_source": src,
                         "kernel": ("em_team_kernel<%s,8,HOT>" if B <= 148 * 32 else "em_kernel<%s>") % (("Env" + env[:-5],) * 1),
                         "kernel_ms_per_launch": kernel_ms, "algorithmic_flops_per_update": F,
                         "algorithmic_bytes_per_update": Bb, "hbm_frac": Bb * rate_gpu / 1e9 / hbm}}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def _mpc_cpu_worker(a):
    from oracle import ref_bench

    return ref_bench.mpc_worker(a)


def mpc_cpu_baseline(cores):
    """PartiallyObservedMpcPolicy.__call__ of the unmodified reference (i2c/policy/mpc.py:156-182), one process per core,
    2 closed-loop roll-outs each (warm start as mpc_quad.py:624-630), 2 untimed + 10 timed control steps per roll-out."""
    import multiprocessing as mp

    if cpu_kind() != "reference":
        return None
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        out = pool.map(_mpc_cpu_worker, [(2, 2, 10, 100 + c) for c in range(cores)])
        wall = time.perf_counter() - t0
    per_solve = max(o[0] / o[1] for o in out)
    return {"value": per_solve * 1e6 / cores, "unit": "us", "per_core_us_per_solve": per_solve * 1e6, "cores": cores,
            "kind": "reference", "seconds": wall,
            "sample": f"{cores} procs x 2 roll-outs x 10 timed control steps; unmodified reference policy/mpc.py with the fp64 "
                      f"quadrotor restatement (Box2D absent); value = us per solve amortised over the {cores} cores"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--problems", type=int, default=4096, help="problems per GPU")
    ap.add_argument("--horizon", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--b-sweep", type=lambda v: [int(x) for x in v.split(",") if x], default=[8192, 16384, 32768, 56832],
                    help="extra batch sizes of the judged workload timed after the headline (N=1 only; '' = none)")
    ap.add_argument("--saturation", type=int, default=65536,
                    help="extra large-batch measurement at N=1 (0 = off); 65536 problems = 2048 tiles on the 1776 resident warps of "
                         "em_ticket_kernel (no wave quantisation)")
    ap.add_argument("--mpc-rollouts", type=int, default=8192, help="roll-outs per GPU of the MPC leg (0 = off)")
    ap.add_argument("--scan-horizon", type=int, default=4096, help="horizon of the parallel-in-time leg at N=1 (0 = off)")
    ap.add_argument("--e2e-steps", type=int, default=100,
                    help="EM iterations of the end-to-end leg (default 100: the whole job of BASELINE configs[2], SURVEY.md 8d; "
                         "0 = --steps); the controllers are read back / gathered once after the last one")
    ap.add_argument("--workload", default="pendulum", choices=["pendulum", "dcp", "quadrotor", "cartpole"],
                    help="pendulum = BASELINE configs[2] (the judged line); dcp / quadrotor = per-GPU shards of configs[3] / [4]")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload != "pendulum":
        run_workload(args, rank, world, local_rank)
    else:
        run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
