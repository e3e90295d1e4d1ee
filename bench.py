#!/usr/bin/env python
"""Headline benchmark: batched pendulum swing-up cubature i2c (BASELINE.json configs[2]:
4096 random initial states x T=200 per B200), metric = problem-timestep updates / s.

  python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port, all host cores)

One "step" = one EM iteration (I2cGraph.learn_msgs: forward + backward + M-step, i2c/i2c.py:1238-1245) over the
whole batch; one problem-timestep update = one cell through one such iteration.  Under torchrun every rank owns an
independent shard of problems (weak scaling, no data-path collective); the only collective is the final NCCL
all_gather of controllers / costs, timed in the end-to-end leg.
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
NCU_DRAM_BYTES_PER_UPDATE = (558.359552e6 + 623.112704e6) / (3 * 4096 * 200)
FP64_NOMINAL_TFLOPS = 37.2  # 148 SM x 64 FMA/clk x 2 x 1.965 GHz


def make_inputs(B, T, seed):
    """Config 3 of SURVEY.md 8(d): x0[b] = [pi,0] + [0.3,0.5] * N(0,I), mu_u[b] = 1e-2 N(0,1)."""
    rng = np.random.default_rng(seed)
    x0 = np.array([np.pi, 0.0]) + np.array([0.3, 0.5]) * rng.normal(size=(B, 2))
    mu_u = 1e-2 * rng.normal(size=(B, T, 1))
    return x0, mu_u


HYPER = dict(Q=np.diag([1.0, 100.0, 1.0]), R=np.diag([2.0]), alpha=100.0, tol=0.0, sig_u=2.0 * np.eye(1))


# ----------------------------------------------------------------------------- CPU reference arm / baseline
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    from oracle import i2c_oracle as O

    seed, Bs, T, steps, warmup = args
    x0, mu_u = make_inputs(Bs, T, seed)
    g = O.make_graph("PendulumKnown", T, HYPER["Q"], HYPER["R"], HYPER["Q"], HYPER["alpha"], HYPER["tol"], mu_u,
                     HYPER["sig_u"], B=Bs, x0=x0)
    for _ in range(warmup):
        g.learn_msgs()
    t0 = time.perf_counter()
    for _ in range(steps):
        g.learn_msgs()
    return time.perf_counter() - t0


def cpu_rate(T, steps, warmup, per_core, cores):
    """Oracle port (batched NumPy restatement of the reference) on `cores` processes, `per_core` problems each."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        times = pool.map(_cpu_worker, [(1000 + i, per_core, T, steps, warmup) for i in range(cores)])
        wall = time.perf_counter() - t0
    dt = max(times)
    return cores * per_core * T * steps / dt, dt, wall


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 64
    T = args.horizon
    rate, dt, wall = cpu_rate(T, args.steps, min(args.warmup, 1), per_core, cores)
    sample = f"{cores} procs x {per_core} problems x T={T}, {args.steps} EM iterations (of the {args.problems}-problem job)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"pendulum swing-up cubature i2c, {args.problems} problems/GPU x T={T} (BASELINE configs[2])",
                   "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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

    # ---- end to end through the public API with HOST buffers (the reference's i2c_run.py loop, :89-98): every
    # step uploads the start-state belief (sys.x0 / sys.sig_x0 are re-read by every sweep), runs one learn_msgs,
    # reads the per-problem cost / alpha scalars and the controller (get_local_linear_policy) back to the host.
    Ke = max(3, min(K, 10))
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()  # noqa: E731
    h_x0, h_s0 = pin(B, 2), pin(B, 2, 2)
    h_x0[:], h_s0[:] = g.x0, g.sig_x0
    h_m = pin(1, B)
    # double-buffered pinned result arrays: the controller copy of step i overlaps the sweep of step i+1
    h_pol = [(pin(B, T, 1, 2), pin(B, T, 1), pin(B, T, 1, 1)) for _ in range(2)]
    h_K, h_k, h_s = h_pol[0]
    L = g.lib

    def e2e_step(i):
        capi.check(L.i2c_set_initial_state_async(g._h, capi.ptr(h_x0), capi.ptr(h_s0)))  # pinned, persistent buffers
        capi.check(L.i2c_run(g._h, 1, capi.PH_LEARN))
        for m in ("alpha", "cost_m"):
            capi.check(L.i2c_get_metric(g._h, capi.METRICS[m], capi.ptr(h_m), 1))
        bK, bk, bs = h_pol[i & 1]
        capi.check(L.i2c_get_policy_async(g._h, capi.ptr(bK), capi.ptr(bk), capi.ptr(bs)))

    def final_gather():
        # the path's only collective: final gather of controllers and costs over NVLink
        from i2c_b200 import dist as idist

        Kd, kd, sd = g.policy_device_tensors()
        cost = torch.from_numpy(np.ascontiguousarray(h_m[0])).to(Kd.device)
        gathered = idist.gather_controllers(Kd, kd, sd, world * B, extra=(cost,))
        assert gathered[0].shape[0] == world * B
        torch.cuda.synchronize(dev)

    e2e_step(0)
    capi.check(L.i2c_copy_wait(g._h))
    if dist is not None:
        final_gather()  # warm-up of the collective (communicator buffers, allocator), like the untimed first step
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        e2e_step(i)
    capi.check(L.i2c_copy_wait(g._h))
    if dist is not None:
        final_gather()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_value = world * B * T * Ke / (e2e_ms * 1e-3)

    # Same loop for a caller that only needs the controllers at the end (EM for Ke iterations, then one read-back): every
    # step still uploads the start-state belief and reads the per-problem cost / alpha back; K, k, sigK cross PCIe once.
    def e2e_step_metrics_only():
        capi.check(L.i2c_set_initial_state_async(g._h, capi.ptr(h_x0), capi.ptr(h_s0)))  # pinned, persistent buffers
        capi.check(L.i2c_run(g._h, 1, capi.PH_LEARN))
        for m in ("alpha", "cost_m"):
            capi.check(L.i2c_get_metric(g._h, capi.METRICS[m], capi.ptr(h_m), 1))

    e2e_step_metrics_only()
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        e2e_step_metrics_only()
    capi.check(L.i2c_get_policy_async(g._h, capi.ptr(h_K), capi.ptr(h_k), capi.ptr(h_s)))
    capi.check(L.i2c_copy_wait(g._h))
    if dist is not None:
        final_gather()
    barrier()
    e2e2_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e2_value = world * B * T * Ke / (e2e2_ms * 1e-3)
    h2d = h_x0.nbytes + h_s0.nbytes
    d2h = h_K.nbytes + h_k.nbytes + h_s.nbytes + 2 * h_m.nbytes

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
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                "what": "per step: H2D start-state belief, one learn_msgs, D2H cost+alpha per problem (sync) and K,k,sigK "
                        "(i2c_get_policy_async into double-buffered pinned arrays: the copy overlaps the next step; all "
                        "copies complete inside the timed region)"
                        + ("; + final NCCL all_gather of controllers" if world > 1 else "")
                        + (f"; process bound to the {numa} CPUs local to its GPU" if numa else "")},
        "e2e_final_readback": {"value": e2e2_value, "unit": UNIT, "steps": Ke,
                               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 2 * h_m.nbytes,
                               "d2h_bytes_once": h_K.nbytes + h_k.nbytes + h_s.nbytes,
                               "what": "as e2e, but the controllers are read back once after the last step (inside the timed "
                                       "region) instead of after every step: the per-step PCIe traffic is the belief upload and "
                                       "the cost / alpha read-back only"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": NCU_DRAM_BYTES_PER_UPDATE * B * T * K if (B == 4096 and T == 200) else None,
                     "traffic_source": "ncu dram bytes/update (profiles/r01i_ncu_full_em_team_kernel_pendulum_4096.txt) x updates "
                                       "per launch; algorithmic bytes per launch = %.4g" % (B_ALG * B * T * K),
                     "peak_source": peak_src,
                     "kernel": "em_team_kernel<EnvPendulum,8>" if B <= 148 * 32 else "em_kernel<EnvPendulum>",
                     "kernel_ms_per_launch": kernel_ms,
                     "algorithmic_bytes_per_update": B_ALG, "algorithmic_flops_per_update": F_ALG,
                     "fp64": {"achieved_tflops": fp64_ach, "peak_measured_tflops": fp64_peak,
                              "frac_of_measured": (fp64_ach / fp64_peak) if fp64_peak else None,
                              "peak_nominal_tflops": FP64_NOMINAL_TFLOPS, "frac_of_nominal": fp64_ach / FP64_NOMINAL_TFLOPS}},
    }
    if world == 1 and args.saturation > B:
        # same kernel at a batch that fills the machine (the 4096-problem headline is bound by the latency of the
        # sequential recursion over t: 128 warps on 592 sub-partitions)
        Bs = args.saturation
        x0s, mu_us = make_inputs(Bs, T, 99)
        del g
        torch.cuda.empty_cache()
        gs = i2c_b200.BatchedI2c("PendulumKnown", Bs, T, HYPER["Q"], HYPER["R"], HYPER["Q"], HYPER["alpha"], HYPER["tol"],
                                 mu_us, HYPER["sig_u"], x0=x0s, device=dev, max_iters=8)
        gs.run(3, capi.PH_LEARN, collect=False)
        gs.run(5, capi.PH_LEARN, collect=False)
        torch.cuda.synchronize(dev)
        ms_s = gs.last_run_ms()
        rate_s = Bs * T * 5 / (ms_s * 1e-3)
        line["saturation"] = {"problems": Bs, "value": rate_s, "unit": UNIT, "ms_per_step": ms_s / 5,
                              "hbm_frac": B_ALG * rate_s / 1e9 / hbm_peak,
                              "hbm_frac_note": "algorithmic 704 B/update; the kernel moves 481 B/update (ncu, packed triangles), "
                                               "i.e. %.2f of the measured copy bandwidth" % (NCU_DRAM_BYTES_PER_UPDATE * rate_s / 1e9 / hbm_peak),
                              "fp64_frac_of_measured": (F_ALG * rate_s / 1e12 / fp64_peak) if fp64_peak else None,
                              "failed_problems": int(np.count_nonzero(gs.status()[0]))}
        del gs
    if mpc is not None:
        line["mpc"] = mpc
    line["em_iterations_per_s"] = {"batch_sweeps_per_s": 1e3 * K / ms, "problem_iterations_per_s": world * B * K / (ms * 1e-3)}
    if world == 1 and args.scan_horizon > 0:
        line["time_parallel"] = scan_leg(args, dev)
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu_it = 24  # ~10-30 s of CPU work per core
        rate, dt, wall = cpu_rate(T, n_cpu_it, 1, 64, cores)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cores} procs x 64 problems x T={T}, {n_cpu_it} EM iterations after 1 warm-up "
                                          f"(oracle/i2c_oracle.py, batched NumPy restatement)", "seconds": wall}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


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


def mpc_leg(args, dev, rank, world, max_over_ranks, barrier):
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
    pol = i2c_b200.BatchedPartiallyObservedMpc(g, mpc_iter, 1e-2 * np.eye(2), z_traj, sig_zeta=sig_zeta)
    pol.set_control(feedforward=False)
    g.calibrate_alpha()
    pol.optimize(25)
    g.calibrate_alpha()
    rng = np.random.default_rng(7 + rank)
    e = g.env
    y0 = np.array([e.x0[0] - 0.8, e.x0[1], e.x0[0] + 0.8, e.x0[1], 0, 0, 0.8, 0.8])  # measure(x0)
    n_warm, n_timed = 3, 20
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
    return {"metric": "us per MPC solve (amortised over the roll-out batch)", "value": ms * 1e3 / (world * B), "unit": "us",
            "higher_is_better": False, "rollouts_per_gpu": B, "batch_latency_ms_per_control_step": ms,
            "solves_per_s": world * B / (ms * 1e-3), "control_steps_timed": n_timed,
            "launches_per_control_step": (g.kernel_launches() - launches0) / n_timed,
            "config": "quadrotor MPC + cubature Kalman filter, T_plan=10, mpc_iter=2, feedback mode (BASELINE configs[4])",
            "failed_rollouts": int(np.count_nonzero(st))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--problems", type=int, default=4096, help="problems per GPU")
    ap.add_argument("--horizon", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--saturation", type=int, default=56832,
                    help="extra large-batch measurement at N=1 (0 = off); default = 148 SMs x 12 warps x 32 problems")
    ap.add_argument("--mpc-rollouts", type=int, default=8192, help="roll-outs per GPU of the MPC leg (0 = off)")
    ap.add_argument("--scan-horizon", type=int, default=4096, help="horizon of the parallel-in-time leg at N=1 (0 = off)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
