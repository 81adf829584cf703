#!/usr/bin/env python
"""bench.py -- env-steps/s (envs x agents) of the Go1.step() hot path on go1gate, num_envs = 4096 per GPU.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA engine behind the mqe VecEnv surface)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the box's host cores

One JSON line on rank 0.  `value` = (global envs x agents x K) / device time, inputs resident in HBM; `e2e` = the same
through the C-ABI call with HOST buffers (H2D of actions, D2H of obs rows + reset flags inside the timed region);
`roofline` = the fused substep kernel (k_substeps), timed live with CUDA events; `cpu_baseline` = the oracle port
on a bounded sample.  See DESIGN.md section "Measurement" for the byte accounting.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TASK = "go1gate"
ENVS_PER_GPU = 4096
BYTES_PER_AGENT_SUBSTEPS = 932           # root 52+52, dof 96+96, actuator hist 192+192, action 48, contact force 204
BYTES_PER_NPC = 104
L2_FLUSH_BYTES = 256 << 20


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_actions(n_envs, a_ctrl, steps, env_offset=0, seed=0):
    """U(-1,1) actions from a generator keyed (seed, step) over the GLOBAL env range so shards draw identical values."""
    out = np.empty((steps, n_envs, a_ctrl, 3), dtype=np.float32)
    for s in range(steps):
        rng = np.random.Generator(np.random.Philox(key=seed * 1_000_003 + s))
        allv = rng.uniform(-1, 1, size=(env_offset + n_envs, a_ctrl, 3)).astype(np.float32)
        out[s] = allv[env_offset:]
    return out


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def time_oracle(n_envs, steps, warmup, threads=None):
    """agent-steps/s of the CPU restatement (oracle/mqe_oracle.c, fp32 build, OpenMP over envs) on `n_envs` envs."""
    import oracle
    from mqe_b200 import scene as S
    from mqe_b200.envs import configs as C
    threads = threads or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    cfg = C.Go1GateCfg()
    cfg.env.num_envs = n_envs
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0, wrapper_action_scale=(2.0, 0.5, 0.5))
    orc = oracle.Oracle(sc, "f32")
    oracle.set_threads(threads, "f32")
    orc.reset()
    acts = synth_actions(n_envs, 2, steps + warmup)
    for s in range(warmup):
        orc.step(acts[s])
    t0 = time.perf_counter()
    for s in range(warmup, warmup + steps):
        orc.step(acts[s])
    dt = time.perf_counter() - t0
    orc.close()
    return n_envs * 2 * steps / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 256
    steps, warmup = min(args.steps, 40), min(args.warmup, 5)
    v, dt, threads = time_oracle(n_sample, steps, warmup)
    line = {
        "impl": "reference", "metric": "env-steps/sec (envs x agents)", "value": v, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{TASK}, 2 Go1 agents, bounded sample of {n_sample} envs per step (of {ENVS_PER_GPU})", "num_envs": n_sample},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n_sample} envs x {steps} policy steps, oracle/mqe_oracle.c fp32, OpenMP; PhysX itself is not runnable (closed Isaac Gym binary)"},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from mqe_b200 import engine as E
    from mqe_b200.dist import StepGather, shard_range
    from mqe_b200.envs import custom_cfg, make_mqe_env
    from types import SimpleNamespace

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_global = args.num_envs * world
    start, stop = shard_range(n_global, rank, world)
    n_local = stop - start
    mode = {"fp32": E.POLICY_FP32, "bf16x3": E.POLICY_BF16X3, "bf16": E.POLICY_BF16}[args.policy]

    eargs = SimpleNamespace(num_envs=n_global, seed=0, headless=True, record_video=False, sim_device=str(dev))
    env, cfg = make_mqe_env(args.task, eargs, custom_cfg(eargs), env_slice=(start, stop), policy_mode=mode)
    base = env.env
    eng = base.engine
    A = base.num_agents
    a_ctrl = base._ctrl_agents
    gather = StepGather(n_global) if world > 1 else None

    K, W = args.steps, max(args.warmup, 3)
    n_act = min(K + W, 64)                                        # distinct action tensors, cycled
    h_actions = synth_actions(n_local, a_ctrl, n_act, env_offset=start)
    if args.actions == "forward":
        h_actions[:] = np.asarray([0.5, 0.0, 0.0], dtype=np.float32)
    d_actions = torch.as_tensor(h_actions, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    obs_rows = eng.tensor(E.BUF_OBS)

    def one_step(i):
        obs, rew, done, info = env.step(d_actions[i % n_act])
        if gather is not None:                                    # the per-step exchange of SURVEY 8(e)
            gather.gather("done", done)
            gather.gather("obs", obs_rows.view(n_local, A, -1))

    env.reset()
    for i in range(W):
        one_step(i)
    torch.cuda.synchronize()

    # ---- timed region: K steps, L2 flushed (untimed) between steps, per-step CUDA events on the launching stream ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = eng.launch_count()
    wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()
        ev0[i].record()
        one_step(W + i)
        ev1[i].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = n_global * A * K / (dev_ms * 1e-3)

    # ---- roofline of the dominant kernel: k_substeps alone, events around its launch ----
    sub_ms = []
    pol_ms = []
    post_ms = []
    R = min(K, 50)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    base._set_scale("wrapper")
    for i in range(R):
        flush.zero_()
        e[0].record(); eng.policy(d_actions[i % n_act].data_ptr())
        e[1].record(); eng.substeps(base.decimation)
        e[2].record(); eng.post_physics()
        e[3].record()
        torch.cuda.synchronize()
        pol_ms.append(e[0].elapsed_time(e[1])); sub_ms.append(e[1].elapsed_time(e[2])); post_ms.append(e[2].elapsed_time(e[3]))
    sub_t = float(np.mean(sub_ms)) * 1e-3
    stats = eng.tensor(E.BUF_STATS).cpu().numpy()
    algo_bytes = n_local * (A * BYTES_PER_AGENT_SUBSTEPS + base.num_npcs * BYTES_PER_NPC)
    peak, peak_src = measured_peaks()
    achieved = algo_bytes / sub_t / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and args.task == TASK and args.num_envs == ENVS_PER_GPU:      # captured on this workload only
        with open(tpath) as f:
            t_ = json.load(f).get("k_substeps", {})
        traffic = t_.get("dram_bytes_read", 0) + t_.get("dram_bytes_write", 0)
    roofline = {"bound": "hbm", "kernel": "k_substeps", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": sub_t * 1e3, "policy_ms": float(np.mean(pol_ms)), "post_ms": float(np.mean(post_ms)),
                "contacts_per_env_substep": float(stats[0]) / max(1, n_local * base.decimation),
                "note": "scalar-fp32 / latency-bound articulated dynamics + PGS: algorithmic bytes per launch are tiny against the time"}
    policy_flops = 2.0 * 1_812_224 * n_local * A
    roofline["policy_tflops"] = policy_flops / (float(np.mean(pol_ms)) * 1e-3) / 1e12
    # the one genuinely dense contraction of the path: the walk policy on the tensor pipe.  bf16x3 issues three bf16
    # MMAs per fp32-equivalent product; the time is the whole policy phase (frame + layer 0 + tail + finish kernels).
    passes = {"fp32": 0, "bf16x3": 3, "bf16": 1}[args.policy]
    tpeak = None
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        with open(ppath) as f:
            tpeak = float(json.load(f).get("bf16_tflops_sustained", 0.0)) or None
    if passes:
        ach = passes * policy_flops / (float(np.mean(pol_ms)) * 1e-3) / 1e12
        roofline["policy"] = {"bound": "tensor", "kernels": "k_policy_frame + k_policy_l0_tc + k_linear_tc x3 + k_body_latent_planes + k_policy_finish",
                              "achieved": ach, "peak": tpeak or 1368.9, "unit": "TFLOP/s", "frac": ach / (tpeak or 1368.9),
                              "peak_source": "measured (bf16 sustained)" if tpeak else "fallback",
                              "note": f"{passes} bf16 MMA passes per fp32-equivalent product; fp32-equivalent rate is policy_tflops"}

    # ---- e2e: through the C-ABI with HOST buffers ----
    h_obs = np.empty((n_local * A, E.OBS_FLOATS), dtype=np.float32)
    h_reset = np.empty(n_local, dtype=np.uint8)
    Ke = min(K, 100)
    for buf in (h_actions, h_obs, h_reset):                       # caller-owned buffers, page-locked once: DMA source / target
        eng.pin_host(buf)
    for i in range(3):
        eng.step_host(h_actions[i % n_act], h_obs, h_reset)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        eng.step_host(h_actions[(3 + i) % n_act], h_obs, h_reset)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {"value": n_global * A * Ke / e2e_s, "unit": "agent-steps/s", "steps": Ke,
           "h2d_bytes_per_step": int(n_local * a_ctrl * 3 * 4), "d2h_bytes_per_step": int(h_obs.nbytes + h_reset.nbytes)}

    line = {
        "metric": "env-steps/sec (envs x agents)", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.task}, {A} Go1 agents, num_envs={args.num_envs} per GPU ({n_global} global), decimation {base.decimation}, "
                               f"dt {cfg.sim.dt}, PGS sweeps {eng.desc.solver_iters}", "policy_arithmetic": args.policy, "actions": "U(-1,1) per step" if args.actions == "uniform" else "fixed (0.5, 0, 0)",
                   "l2": "flushed between timed steps (256 MiB memset, untimed; per-step CUDA events summed)",
                   "sharding": "contiguous env blocks per rank; NCCL all_gather of obs rows + done per step" if world > 1 else "single GPU",
                   "env_steps_per_s": value / A, "physics_substeps_per_s": value / A * base.decimation},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": launches / K,
        "wall_s_timed_region": wall, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, threads = time_oracle(256, 20, 2)
        line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                "sample": "256 envs x 20 policy steps of the same workload, oracle/mqe_oracle.c fp32 + OpenMP (CPU restatement, not PhysX)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--task", type=str, default=TASK)
    ap.add_argument("--num-envs", type=int, default=ENVS_PER_GPU, help="environments per GPU")
    ap.add_argument("--policy", type=str, default=os.environ.get("MQE_BENCH_POLICY", "bf16x3"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--actions", type=str, default="uniform", choices=["uniform", "forward"],
                    help="SURVEY 8(d): U(-1,1) per step (default), or the fixed pattern a = (0.5, 0, 0): every robot walks straight into the gate (contact-heavy)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
