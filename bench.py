#!/usr/bin/env python
"""bench.py -- env-steps/s (envs x agents) of the Go1.step() hot path on go1gate, num_envs = 4096 per GPU (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA engine behind the mqe VecEnv surface)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the box's host cores

One JSON line on rank 0.  `value` = (global envs x agents x K) / device time, inputs resident in HBM; `e2e` = the same through the
C-ABI call with HOST buffers (H2D of actions, D2H of the step result inside the timed region); `roofline` = the fused substep kernel
(k_substeps), timed live with CUDA events (HBM fraction as the contract asks, plus the issue-slot fraction that actually bounds it);
`cpu_baseline` = the oracle port on a bounded sample; `configs` = the other BASELINE.json configurations (C3, C4, C5's per-GPU share;
C5 itself under --gpus 8) measured the same way.

Steady state, not episode phase: every env gets a random episode phase and the whole batch is rolled forward one full episode before
anything is timed (`config.pre_roll_steps`), so resets, falls and contacts are spread the way they are in a long training run and the
result does not depend on --steps.  See DESIGN.md section "Measurement" for the byte accounting.
"""
from __future__ import annotations

import argparse
import json
import os
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
SUB_CONFIGS = [("go1sheep-hard", 4096, "C3"), ("go1seesaw", 8192, "C4"), ("go1football-defender", 4096, "C5 per-GPU share")]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1368.9, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled IN PROCESS through NVML every ~0.5 ms while the timed region runs (the recipe's clocks line;
    `nvidia-smi -lms 100` cannot see a 7 ms region).  Falls back to one nvidia-smi query if NVML is unavailable."""

    def __init__(self, index=0, period_s=0.0005):
        self.index, self.period, self.rows, self.stop_flag, self.thread, self.h = index, period_s, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.h = None

    def _run(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:  # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((time.perf_counter(), float(sm), int(rs)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self, t0=None, t1=None):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=2)
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)] or self.rows
        sm = [r[1] for r in rows]
        bits = 0
        for r in rows:
            bits |= r[2]
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": [n for n, b in names.items() if bits & b],
                "samples": len(sm), "sampler": "nvml in-process, %.1f ms period" % (1e3 * self.period)}


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
def desynchronise(base, torch, seed):
    """Random episode phase per env (keyed by the GLOBAL env id so shards agree), then the caller pre-rolls one full episode: afterwards
    time-outs, falls and contacts are spread uniformly over the batch, as in a long training run."""
    n0 = int(base.scene.desc.env_id_offset)
    g = np.random.Generator(np.random.Philox(key=seed * 7_919 + 17))
    phase = g.integers(0, base.max_episode_length, size=n0 + base.num_envs)[n0:]
    base.episode_length_buf.copy_(torch.as_tensor(phase, device=base.device, dtype=base.episode_length_buf.dtype))


def measure(args, task, n_per_gpu, K, W, ctx, headline):
    """One workload: returns the result dict (value, e2e, roofline ...).  ctx: torch / dist handles and rank info."""
    torch, dist, world, rank, dev = ctx["torch"], ctx["dist"], ctx["world"], ctx["rank"], ctx["dev"]
    from types import SimpleNamespace

    from mqe_b200 import engine as E
    from mqe_b200.dist import PeerStepExchange, StepGather, shard_range
    from mqe_b200.envs import custom_cfg, make_mqe_env
    from mqe_b200.openrl_adapter import mqe_openrl_wrapper

    n_global = n_per_gpu * world
    start, stop = shard_range(n_global, rank, world)
    n_local = stop - start
    mode = {"fp32": E.POLICY_FP32, "bf16x3": E.POLICY_BF16X3, "bf16": E.POLICY_BF16}[args.policy]
    eargs = SimpleNamespace(num_envs=n_global, seed=0, headless=True, record_video=False, sim_device=str(dev))
    env, cfg = make_mqe_env(task, eargs, custom_cfg(eargs), env_slice=(start, stop), policy_mode=mode)
    base = env.env
    eng = base.engine
    A, a_ctrl = base.num_agents, base._ctrl_agents
    env.reset()                                                   # task wrappers switch to the fused gather here
    fused = bool(getattr(env, "_fused", False))

    # ---- multi-GPU: the per-step exchange of what the learner reads, inside the step graph (csrc/gather.cu) ----
    exchange, gather, sharding = None, None, "single GPU"
    if world > 1:
        if args.exchange == "p2p" and n_local % 16 == 0 and len({shard_range(n_global, r, world)[1] - shard_range(n_global, r, world)[0] for r in range(world)}) == 1:
            exchange = PeerStepExchange(eng)
            sharding = "contiguous env blocks per rank; packed step result (wrapper obs | reward | done) stored into every peer's buffer over NVLink by the last kernel of the step graph"
        else:
            gather = StepGather(world)                            # one packed row per rank
            sharding = "contiguous env blocks per rank; ONE NCCL all_gather of the packed step result per step"
        env.reset()                                               # first exchange: global reset observation

    n_act = 64                                                    # distinct action tensors, cycled (independent of K and W: the pre-rolled state must not depend on --steps)
    h_actions = synth_actions(n_local, a_ctrl, n_act, env_offset=start)
    if args.actions == "forward":
        h_actions[:] = np.asarray([0.5, 0.0, 0.0], dtype=np.float32)
    d_actions = torch.as_tensor(h_actions, device=dev)
    flush = ctx["flush"]
    result_all = eng.tensor(E.BUF_STEP_RESULT)

    def one_step(i):
        env.step(d_actions[i % n_act])
        if gather is not None:
            gather.gather("result", result_all[eng.result_parity()].view(1, -1))

    # ---- steady state: random episode phases, then one full episode of pre-roll (untimed) ----
    pre = 0 if (args.no_preroll or ctx.get("no_preroll")) else (args.preroll_episodes if headline else 1) * int(base.max_episode_length) + 1
    if pre:
        desynchronise(base, torch, seed=0)
    for i in range(pre):
        if i >= pre - 40:                                          # the last pre-roll steps already run in the timed region's regime (L2 flushed every step)
            flush.zero_()
        one_step(i)
    # the clock sampler thread is spun up BEFORE the warm-up steps: it is then polling when even a 7 ms timed region starts, and the GPU
    # never sits idle between warm-up and the timed region (a 20 ms pause here cost the first ~15 timed steps 10 % in an earlier version)
    sampler = ClockSampler(ctx["local_rank"]) if (headline and rank == 0) else None
    if sampler:
        sampler.start()
    for i in range(W):
        flush.zero_()
        one_step(pre + i)
    torch.cuda.synchronize()

    # ---- timed region: K steps, L2 flushed (untimed) between steps, per-step CUDA events on the launching stream ----
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
        one_step(pre + W + i)
        ev1[i].record()
    torch.cuda.synchronize()
    wall1 = time.perf_counter()
    wall = wall1 - wall0
    clocks = sampler.stop(wall0, wall1) if sampler else None
    if world > 1:
        dist.barrier()
    launches = eng.launch_count() - launches0
    per_step = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    dev_ms = sum(per_step)
    q = max(1, K // 4)
    drift = {"first_quarter_ms": float(np.mean(per_step[:q])), "last_quarter_ms": float(np.mean(per_step[-q:])),
             "p50_ms": float(np.median(per_step)), "p95_ms": float(np.quantile(per_step, 0.95)), "max_ms": float(np.max(per_step)),
             "first_32_ms": [round(float(x), 4) for x in per_step[:32]]}
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = n_global * A * K / (dev_ms * 1e-3)
    timed_out = bool(exchange.timed_out()) if exchange is not None else False

    # ---- kernel legs in the same (steady-state) regime: policy / substeps / post separately, CUDA events around each ----
    sub_ms, pol_ms, post_ms = [], [], []
    R = max(10, min(K, 50))
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    base._set_scale("wrapper")
    contacts = 0.0
    for i in range(R):
        flush.zero_()
        e[0].record(); eng.policy(d_actions[i % n_act].data_ptr())
        e[1].record(); eng.substeps(base.decimation)
        e[2].record(); eng.post_physics()
        e[3].record()
        torch.cuda.synchronize()
        pol_ms.append(e[0].elapsed_time(e[1])); sub_ms.append(e[1].elapsed_time(e[2])); post_ms.append(e[2].elapsed_time(e[3]))
        contacts += float(eng.tensor(E.BUF_STATS)[0].item())
    # the same stages INSIDE the step graph (event marks between them; the marks cost the launch overlap between stages, ~0.01-0.02 ms)
    eng.stage_timing(True)
    for i in range(3):
        one_step(i)
    st_rows = []
    for i in range(R):
        flush.zero_()
        one_step(3 + i)
        st_rows.append(list(eng.stage_ms().values()))
    eng.stage_timing(False)
    st_mean = np.mean(np.asarray(st_rows), axis=0)
    in_step = {"policy_ms": float(st_mean[0]), "physics_and_bookkeeping_ms": float(st_mean[1]), "separate_bookkeeping_gather_exchange_ms": float(st_mean[2]),
               "background_join_ms": float(st_mean[3]),
               "note": "mqe_sim_stage_ms: event marks recorded as nodes of the step graph; post-physics bookkeeping runs in the epilogue of k_substeps "
                       "(fused), the 29-frame layer-0 pass of the next step runs behind the physics and is joined at the end"}
    sub_t = float(np.mean(sub_ms)) * 1e-3
    algo_bytes = n_local * (A * BYTES_PER_AGENT_SUBSTEPS + base.num_npcs * BYTES_PER_NPC)
    peaks, peak_src = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    achieved = algo_bytes / sub_t / 1e9
    traffic, inst = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and task == TASK and n_per_gpu == ENVS_PER_GPU:      # captured on this workload only
        with open(tpath) as f:
            t_ = json.load(f).get("k_substeps", {})
        traffic = t_.get("dram_bytes_read", 0) + t_.get("dram_bytes_write", 0)
        inst = t_.get("warp_inst_executed")
    roofline = {"bound": "hbm", "kernel": "k_substeps", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": sub_t * 1e3, "policy_ms": float(np.mean(pol_ms)), "post_ms": float(np.mean(post_ms)),
                "policy_ms_on_critical_path": in_step["policy_ms"], "in_step": in_step,
                "policy_note": "kernel_ms / policy_ms / post_ms are the three stand-alone C-ABI calls (mqe_sim_substeps, mqe_sim_policy = frame + full 30-frame "
                               "layer 0 + fused tail, mqe_sim_post_physics); inside mqe_sim_step 29 of the 30 frames were contracted behind the previous step's "
                               "physics and the bookkeeping is fused into k_substeps, so the step pays `in_step` instead",
                "contacts_per_env_substep": contacts / max(1, R * n_local * base.decimation),
                "regime": f"steady state (random episode phases, {pre} pre-roll steps), {R} launches after the timed region",
                "note": "scalar-fp32 articulated dynamics + PGS: bound by instruction issue / latency, not by HBM -- see `issue`"}
    if inst:
        sm_hz = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
        floor_s = float(inst) / (148 * 4 * sm_hz)                # one warp instruction per scheduler per clock
        roofline["issue"] = {"warp_inst_per_launch": inst, "source": "ncu smsp__inst_executed.sum (profiles/ncu_traffic.json)",
                             "floor_ms": floor_s * 1e3, "frac": floor_s / sub_t, "peak": "148 SM x 4 schedulers x %.0f MHz" % (sm_hz / 1e6)}
    policy_flops = 2.0 * 1_812_224 * n_local * A
    roofline["policy_tflops"] = policy_flops / (float(np.mean(pol_ms)) * 1e-3) / 1e12
    passes = {"fp32": 0, "bf16x3": 3, "bf16": 1}[args.policy]
    tpeak = float(peaks.get("bf16_tflops_sustained", 0.0)) or 1368.9
    if passes:
        ach = passes * policy_flops / (float(np.mean(pol_ms)) * 1e-3) / 1e12
        roofline["policy"] = {"bound": "tensor", "kernels": "policy phase: frame + tcgen05 layer 0 + fused tail + finish",
                              "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                              "peak_source": "measured (bf16 sustained)" if peak_src == "measured" else "fallback",
                              "note": f"{passes} bf16 MMA passes per fp32-equivalent product; fp32-equivalent rate is policy_tflops"}

    # ---- e2e: HOST buffers in, HOST buffers out, through the call a user makes ----
    Ke = max(10, min(K, 100))
    if fused:
        # the reference-facing L5 path: mqe_openrl_wrapper.step(numpy actions) -> numpy obs / rewards / dones (openrl_ws/utils.py:53-67);
        # one pinned H2D, the step (incl. the peer exchange when sharded), ONE packed D2H (mqe_sim_step_host_result)
        ad = mqe_openrl_wrapper(env)
        acts2 = 2.0 * h_actions                                   # the adapter halves its input (utils.py:55)
        for i in range(3):
            ad.step(acts2[i % n_act])
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            obs, rew, dones, _ = ad.step(acts2[(3 + i) % n_act])
        e2e_s = time.perf_counter() - t0
        d2h = int(ad._host["L"].total_bytes)
        api = "mqe_openrl_wrapper.step(np actions) -> np obs, rewards, dones (fused wrapper, mqe_sim_step_host_result)"
    else:
        # go1gate: the shipped wrapper returns (0, 0, done) (go1_gate_wrapper.py:155), so the learner-facing result is the done flags;
        # to keep the observation traffic in the number, e2e copies the full obs_buf rows [N*A][71] back as well (mqe_sim_step_host)
        h_obs = np.empty((n_local * A, E.OBS_FLOATS), dtype=np.float32)
        h_reset = np.empty(n_local, dtype=np.uint8)
        for buf in (h_actions, h_obs, h_reset):
            eng.pin_host(buf)
        for i in range(3):
            eng.step_host(h_actions[i % n_act], h_obs, h_reset)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            eng.step_host(h_actions[(3 + i) % n_act], h_obs, h_reset)
        e2e_s = time.perf_counter() - t0
        d2h = int(h_obs.nbytes + h_reset.nbytes)
        api = "mqe_sim_step_host (C ABI): pinned H2D of actions, step, D2H of obs_buf rows + done flags"
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {"value": n_global * A * Ke / e2e_s, "unit": "agent-steps/s", "steps": Ke, "h2d_bytes_per_step": int(n_local * a_ctrl * 3 * 4),
           "d2h_bytes_per_step": d2h, "api": api, "includes_exchange": world > 1}

    out = {
        "value": value, "ms_per_step": dev_ms / K, "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": launches / K,
        "wall_s_timed_region": wall, "roofline": roofline, "clocks": clocks, "per_step_ms": drift,
        "config": {"workload": f"{task}, {A} Go1 agents" + (f" + {base.num_npcs} NPC" if base.num_npcs else "") +
                               f", num_envs={n_per_gpu} per GPU ({n_global} global), decimation {base.decimation}, dt {cfg.sim.dt}, PGS sweeps {eng.desc.solver_iters}",
                   "policy_arithmetic": args.policy, "actions": "U(-1,1) per step" if args.actions == "uniform" else "fixed (0.5, 0, 0)",
                   "l2": "flushed between timed steps (256 MiB memset, untimed; per-step CUDA events summed)",
                   "pre_roll_steps": pre, "episode_phase": "random per env (steady state)" if pre else "all envs reset together at step 0",
                   "sharding": sharding, "fused_wrapper_gather": fused,
                   "env_steps_per_s": value / A, "physics_substeps_per_s": value / A * base.decimation},
    }
    if timed_out:
        out["exchange_timeout"] = True
    env.close()
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = {"torch": torch, "dist": dist, "world": world, "rank": rank, "local_rank": local_rank, "dev": dev,
           "flush": torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)}
    K, W = args.steps, max(args.warmup, 3)
    main = measure(args, args.task, args.num_envs, K, W, ctx, headline=True)
    line = {
        "metric": "env-steps/sec (envs x agents)", "value": main["value"], "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.policy == "fp32" else f"f32 physics / {args.policy} policy (bf16 hi+lo operands, fp32 accumulate)" if args.policy == "bf16x3" else "f32 physics / bf16 policy",
        "data": "synthetic", "config": main["config"], "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"],
        "launches_per_step": main["launches_per_step"], "wall_s_timed_region": main["wall_s_timed_region"], "per_step_ms": main["per_step_ms"],
        "roofline": main["roofline"],
    }
    if "exchange_timeout" in main:
        line["exchange_timeout"] = True
    # ---- the other BASELINE.json configurations, measured the same way (sub-lines; the headline above is config 2) ----
    if not args.no_sublines and args.task == TASK:
        subs = []
        Ks = max(10, min(K, 50))
        todo = SUB_CONFIGS if world == 1 else ([("go1football-defender", 4096, f"C5: 3 Go1 + ball, num_envs={4096 * world} sharded across {world} GPUs")] if world == 8 or args.c5 else [])
        for task, n, label in todo:
            r = measure(args, task, n, Ks, 3, ctx, headline=False)
            subs.append({"baseline_config": label, "workload": r["config"]["workload"], "value": r["value"], "unit": "agent-steps/s", "ms_per_step": r["ms_per_step"],
                         "steps": Ks, "e2e": r["e2e"], "kernel_ms": r["roofline"]["kernel_ms"], "policy_ms": r["roofline"]["policy_ms"], "post_ms": r["roofline"]["post_ms"],
                         "roofline_frac_hbm": r["roofline"]["frac"], "contacts_per_env_substep": r["roofline"]["contacts_per_env_substep"],
                         "pre_roll_steps": r["config"]["pre_roll_steps"], "launches_per_step": r["launches_per_step"]})
        line["configs"] = subs
    if world == 1 and not args.no_sublines and not args.no_preroll and args.task == TASK:
        # round 1 timed steps 5..25 after a synchronised reset (every robot still settling, few contacts); kept as a side number so that
        # BENCH_r01 and this round can be compared like for like -- the headline above is the steady state
        r = measure(args, args.task, args.num_envs, 20, 5, dict(ctx, no_preroll=True), headline=False)
        line["early_episode"] = {"value": r["value"], "ms_per_step": r["ms_per_step"], "kernel_ms": r["roofline"]["kernel_ms"],
                                 "contacts_per_env_substep": r["roofline"]["contacts_per_env_substep"],
                                 "note": "round-1 regime (all envs at episode steps 5..25); comparison only, not the headline"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, threads = time_oracle(256, 20, 2)
        line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                "sample": "256 envs x 20 policy steps of the same workload, oracle/mqe_oracle.c fp32 + OpenMP (CPU restatement, not PhysX)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--task", type=str, default=TASK)
    ap.add_argument("--num-envs", type=int, default=ENVS_PER_GPU, help="environments per GPU")
    ap.add_argument("--policy", type=str, default=os.environ.get("MQE_BENCH_POLICY", "bf16x3"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sublines", action="store_true", help="skip the C3 / C4 / C5 sub-lines")
    ap.add_argument("--no-preroll", action="store_true", help="time from a synchronised reset instead of the steady state (round-1 behaviour)")
    ap.add_argument("--preroll-episodes", type=int, default=3, help="episodes rolled forward (untimed) after the phase randomisation, headline workload")
    ap.add_argument("--c5", action="store_true", help="under torchrun with fewer than 8 ranks: still add the sharded football-defender sub-line")
    ap.add_argument("--exchange", type=str, default=os.environ.get("MQE_EXCHANGE", "p2p"), choices=["p2p", "nccl"],
                    help="multi-GPU per-step exchange: peer-memory stores inside the step graph (default) or one NCCL all_gather")
    ap.add_argument("--actions", type=str, default="uniform", choices=["uniform", "forward"],
                    help="SURVEY 8(d): U(-1,1) per step (default), or the fixed pattern a = (0.5, 0, 0): every robot walks straight into the gate (contact-heavy)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
