#!/usr/bin/env python
"""bench.py -- SkyJo env-steps/sec of the fused step + mask + observe kernel on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # N=1: this process
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one ITERATION of a rollout loop: 64 launches of the fused kernel -- each launch is one
env-step (SkyjoGame.act + the next agent's observation and action mask, i.e. one iteration of reference
rlskyjo/game/sample_game.py:10-21) for EVERY env of the batch, with the uniform legal policy drawn in-kernel and
auto-reset on -- followed by the one collective of the path, the all-reduce of the 32-entry statistics vector, on a
side stream.  `--steps K` therefore times K x 64 launches (K = 20 -> about 60 ms), `ms_per_step` is per iteration,
and `value` stays in env-steps/s.
Auto-reset runs in the phase-locked "next_step" mode by default (include/skyjo_b200.h SKYJO_RESET_NEXT_STEP): the
lockstep slot an env spends on its reset is NOT an env-step and is not counted -- every rate below divides the act()
transitions counted by the kernels' statistics vector (SKYJO_STAT_STEPS), not launches x envs.
Workload = BASELINE.json configs[1]: 4-player SkyJo, 2^20 lockstep envs per GPU, direct observations (D = 67).
Envs shard over GPUs by global env id (weak scaling: 2^20 per GPU).

Printed JSON (rank 0): value = whole-job env-steps/s with state resident in HBM; e2e = the same metric through the
host-buffer C-ABI entry (skyjo_step_host: pinned actions in, obs / mask / agent / done / reward out, copies inside
the timed region); roofline = algorithmic bytes of one launch (36N + 119 B per env-step with the
policy fused) / the step kernel's mean device time, against the measured HBM copy bandwidth; cpu_baseline = the C
oracle port of the reference's loop on this box's host cores, with the unmodified Python reference (baseline/_ref,
BASELINE config 1: N=2, 10 000 games, multiprocessing.Pool) beside it; configs.c3 / configs.c5 = BASELINE configs 3
(8 players, 2^24 envs on one GPU) and 5 (2^23 envs per GPU, 64 M on 8 GPUs) measured by the same rules in the same
run, each with its own roofline.

`--impl reference` times the CPU implementation alone (all host threads) on the same workload definition: the C
oracle port (the reference is pure Python, nothing compiles into oracle/_ref), and reports the Python reference's
own rate beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SkyJo env-steps/sec"
UNIT = "env-steps/s"
L2_MB = 126.0
EP_LEN = {1: 46, 2: 76, 3: 106, 4: 133, 8: 240, 12: 341}     # mean act() calls per game under the uniform policy


def algorithmic_bytes_per_step(N, indirect, action_bytes=0):
    # read 24N+24, write 48, outputs D+28, + action bytes
    return (24 * N + 131 + action_bytes) if indirect else (36 * N + 119 + action_bytes)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(key):
    """Per-launch DRAM bytes of the step kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:  # noqa: BLE001
            return None
    return None


# ---- CPU baselines -------------------------------------------------------------------------------
_POOL = None


def cpu_rollout(N, indirect, games_per_thread, threads, seed=0, env0=0):
    """All threads play `games_per_thread` full games each on the C oracle port; returns (env-steps, seconds)."""
    global _POOL
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O
    O.lib()
    if _POOL is None or _POOL._max_workers != threads:
        _POOL = ThreadPoolExecutor(max_workers=threads)  # ctypes releases the GIL

    def work(i):
        return O.rollout(N, 2.0, indirect, seed, env0 + i * games_per_thread, games_per_thread)[0]

    t0 = time.perf_counter()
    steps = sum(_POOL.map(work, range(threads)))
    return steps, time.perf_counter() - t0


def python_reference(games=10000, timeout=600):
    """BASELINE config 1 on the UNMODIFIED Python reference (baseline/_ref), in a fresh interpreter so that its
    multiprocessing pool never meets this process's CUDA context: tools/python_reference_rate.py."""
    tool = os.path.join(ROOT, "tools", "python_reference_rate.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "rlskyjo")):
        return {"unavailable": "baseline/_ref/rlskyjo is missing (created by __graft_entry__.build() where "
                               "/root/reference is mounted)"}
    try:
        r = subprocess.run([sys.executable, tool, "--json", "--games", str(games), "--ref",
                            os.path.join(ROOT, "baseline", "_ref")], capture_output=True, text=True, timeout=timeout)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            return {"unavailable": f"rc {r.returncode}: {(r.stderr or r.stdout)[-300:]}"}
        return json.loads(line[-1])
    except Exception as ex:  # noqa: BLE001
        return {"unavailable": repr(ex)[:300]}


def cpu_baseline_record(N, indirect, budget_s, with_python=True):
    threads = os.cpu_count() or 1
    games = max(2, int(budget_s * 1.0e6 / EP_LEN.get(N, 30 * N) / threads))
    cpu_rollout(N, indirect, max(2, games // 10), threads)
    s, dt = cpu_rollout(N, indirect, games, threads)
    rec = {"value": s / dt, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{games} full games per thread x {threads} threads ({s} env-steps), C oracle port of "
                     "sample_game.py:10-21 with the same Philox deal and uniform legal policy"}
    if with_python:
        rec["python_reference"] = python_reference()
    return rec


# ---- clocks ------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.sm_max = None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---- arms ----------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    N, ind = args.players, args.indirect
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    # bounded sample per step so that the whole run stays near args.cpu_budget seconds
    games = int(args.cpu_budget * 1.0e6 / EP_LEN.get(N, 30 * N) / max(total, 1))
    games = max(2, min(games, 20000))
    for w in range(args.warmup):
        cpu_rollout(N, ind, games, threads, env0=w * threads * games)
    steps = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        s, _ = cpu_rollout(N, ind, games, threads, env0=(args.warmup + k) * threads * games)
        steps += s
    dt = time.perf_counter() - t0
    value = steps / dt
    sample = f"{games} full games per thread per step x {threads} threads, C oracle port, Philox deal + uniform legal policy"
    cpu = {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    if not args.no_python_reference and args.gpus == 1:
        cpu["python_reference"] = python_reference()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": workload_config(args.players, args.envs, args.indirect, args.reset, world, args),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cores": threads,
    }
    print(json.dumps(line), flush=True)


def workload_config(N, B, indirect, reset, world, args):
    D = 31 if indirect else 19 + 12 * N
    # per env-step: 1+N planes of 16 B read, plane 0 + (every other step) one row written, outputs
    per_step_mb = B * (16 * (1 + N) + 16 + 8 + D + 28) / 1e6
    return {
        "workload": f"{N}-player SkyJo, {B} lockstep envs per GPU, uniform legal policy in-kernel, "
                    f"fused step+mask+observe, {'indirect' if indirect else 'direct'} obs D={D}, "
                    f"auto-reset ({reset})",
        "step": f"one iteration = {args.launches_per_step} env-step launches + one statistics all-reduce (side stream)",
        "env_steps_per_step": args.launches_per_step,
        "reset": reset + (" (phase-locked: the slot an env spends on its reset is not an env-step and is not "
                          "counted; value = counted act() transitions / time)" if reset == "next_step" else ""),
        "num_players": N, "envs_per_gpu": B, "global_envs": B * world, "obs_len": D,
        "parallelism": f"env-sharded x{world}, no data-path collective; per GPU the batch is stepped as 4 "
                       "independent env ranges on 4 CUDA streams; an iteration's launches are replayed from one CUDA graph "
                       "(graph_replays counts them)",
        "l2": f"no flush: ~{per_step_mb:.0f} MB touched per launch vs {L2_MB:.0f} MB L2 (inputs larger than L2)",
        "preroll_steps": args.preroll,
    }


class Ctx:
    """torch / distributed plumbing shared by the measurements of one run"""

    def __init__(self, args, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        # the per-iteration statistics all-reduce (the only collective): the library's own ncclAllReduce on a
        # communicator created with the NCCL torch loaded; torch.distributed.all_reduce if that cannot be created
        self.comm = None
        if world > 1:
            try:
                from skyjo_rl_b200.nccl import StatsComm
                self.comm = StatsComm(self.dev)
            except Exception as ex:  # noqa: BLE001
                print(f"[bench] in-library NCCL communicator unavailable ({ex}); using torch.distributed", file=sys.stderr)
            ok = torch.tensor([1 if self.comm is not None else 0], device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self.comm = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world > 1:
            self.dist.all_reduce(v)
        return v


def measure_step_loop(cx, N, B, indirect, reset, K, W, preroll, seed, profile_steps=512, keep_env=False):
    """The timed loop of one workload: K iterations of (L step launches, one side-stream statistics all-reduce),
    CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks; then the step kernel's
    own launch duration for the roofline.  Returns (record, env or None)."""
    torch, dist = cx.torch, cx.dist
    from skyjo_rl_b200 import BatchedSkyjoEnv, _lib
    L, world, dev = cx.args.launches_per_step, cx.world, cx.dev
    i_steps = _lib.STAT_NAMES.index("steps")
    env = BatchedSkyjoEnv(num_envs=B, num_players=N, observe_other_player_indirect=indirect,
                          device=dev, seed=seed, first_global_env_id=cx.rank * B, auto_reset=reset)
    env.reset()
    use_lib = world == 1 or cx.comm is not None
    last = [None]

    def iteration():
        env.step_random(L)
        if use_lib:
            last[0] = env.stats_allreduce_async(cx.comm)      # local reduction in stream, the collective off it
        else:
            last[0] = env.stats_tensor()
            dist.all_reduce(last[0])

    done = 0
    while done < preroll:            # desynchronise the episodes: steady-state mix of game stages
        env.step_random(min(L, preroll - done))
        done += L
    for _ in range(W):
        iteration()
    env.stats_allreduce_wait()
    env.clear_stats()
    l0, g0 = env.launch_count, env.graph_replay_count
    sampler = ClockSampler(cx.local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.barrier()
    sampler.start()
    ev0.record()
    for _ in range(K):
        iteration()
    env.stats_allreduce_wait()       # the timed region ends when the last all-reduce has landed
    ev1.record()
    cx.barrier()
    clocks = sampler.stop()
    ms = cx.max_over_ranks(ev0.elapsed_time(ev1))
    launches = env.launch_count - l0
    replays = env.graph_replay_count - g0
    reduced = last[0].clone()        # the library's all-reduced vector of the last iteration = the whole region
    check = cx.sum_over_ranks(env.stats_tensor())
    assert torch.equal(reduced, check), "side-stream statistics all-reduce disagrees with torch.distributed.all_reduce"
    counted = int(check[i_steps].item())
    stats = dict(zip(_lib.STAT_NAMES, check.tolist()))
    value = counted / (ms * 1e-3)

    # roofline of the dominant kernel: mean device time of a step launch (event pair per window of back-to-back launches)
    env.clear_stats()
    prof = env.step_random_profile(min(K * L, profile_steps))
    step_us = 1e3 * prof["step_ms"] / max(prof["step_launches"], 1)
    # algorithmic bytes of one launch = bytes per env-step x env-steps the launch plays (envs in a reset slot move
    # state but play no step: they are not credited)
    bps = algorithmic_bytes_per_step(N, indirect)
    alg = bps * int(env.stats_tensor()[i_steps].item()) / max(prof["step_launches"], 1)
    peak, peak_src = hbm_peak()
    achieved = alg / (step_us * 1e-6) / 1e9
    key = f"step_N{N}_{'indirect' if indirect else 'direct'}_B{B}"
    traffic = committed_traffic(key)
    loop_gbs = bps * (counted / world) / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "skyjo::step_kernel", "kernel_us": step_us,
                # the int8 planes move fewer bytes than SURVEY 8(d)'s accounting (80 B of planes read per 4-player
                # env-step instead of 120, ...): the DRAM bytes ncu measured per launch over the same kernel time
                "traffic_frac": (traffic / (step_us * 1e-6) / 1e9 / peak) if traffic else None,
                "algorithmic_bytes_per_env_step": bps, "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                "deal_kernel_share": prof["deal_ms"] / max(prof["deal_ms"] + prof["step_ms"], 1e-9),
                "note": "kernel_us = mean duration of a full-batch step launch on one stream: CUDA events bracket every "
                        "window of back-to-back launches between two refill deals (32 at four players) (no event between launches; the "
                        "deals, in stream order, are timed by their own pairs); the timed loop behind `value` steps "
                        "the batch as 4 env ranges on 4 streams, so that one range's launch ramp / tail is covered by "
                        "the others",
                "loop_frac": loop_gbs / peak,
                "loop_frac_note": "same algorithmic bytes / whole timed loop (refill deals and statistics included): "
                                  "a lower bound of the step kernel's fraction inside the loop"}
    env.check()
    rec = {"value": value, "unit": UNIT, "steps": K, "warmup": W, "ms_per_step": ms / K,
           "us_per_launch": 1e3 * ms / (K * L), "timed_region_s": ms * 1e-3, "env_steps_per_step": L,
           "gpu_launches": launches, "graph_replays": replays, "counted_env_steps": counted, "counted_frac": counted / float(B * world * K * L),
           "roofline": roofline, "clocks": clocks,
           "episode_stats": {k: stats[k] for k in ("episodes", "episode_steps", "refunds", "reshuffles", "steps")}}
    if not keep_env:
        env.close()
        del env
        torch.cuda.empty_cache()
        env = None
    return rec, env


def run_ours(args, rank, local_rank, world):
    cx = Ctx(args, rank, local_rank, world)
    torch, dist, dev = cx.torch, cx.dist, cx.dev
    from skyjo_rl_b200 import BatchedSkyjoEnv, _lib
    N, B, K, W, L = args.players, args.envs, args.steps, args.warmup, args.launches_per_step

    main, env = measure_step_loop(cx, N, B, args.indirect, args.reset, K, W, args.preroll, args.seed, keep_env=True)
    peak = main["roofline"]["peak"]

    # the same env-steps as multi-step launches (rollout_kernel: state in registers across 8 steps, every
    # step's obs / mask / agent / done stored into slice t of time-major rollout tensors)
    rollout = None
    if args.rollout_steps > 0:
        D = env.obs_len
        T = min(args.rollout_steps, int(8e9 // (B * (D + 28))) // 8 * 8)
        if T >= 8:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ro = env.rollout_random(T)
            reps = max(1, min(K * L, 1024) // T)
            for _ in range(2):
                env.rollout_random(T, ro)
            env.clear_stats()
            cx.barrier()
            ev0.record()
            for _ in range(reps):
                env.rollout_random(T, ro)
                env.stats_allreduce_async(cx.comm) if (world == 1 or cx.comm is not None) else None
            env.stats_allreduce_wait()
            ev1.record()
            cx.barrier()
            rms = cx.max_over_ranks(ev0.elapsed_time(ev1)) / (reps * T)
            rcounted = int(cx.sum_over_ranks(env.stats_tensor())[_lib.STAT_NAMES.index("steps")].item())
            env.profile_begin()
            for _ in range(reps):
                env.rollout_random(T, ro)
            rp = env.profile_end()
            env.check()
            rollout = {"value": rcounted / (rms * 1e-3 * reps * T), "unit": UNIT, "ms_per_env_step": rms,
                       "kernel": "skyjo::rollout_kernel", "env_steps_per_launch": 8, "rollout_len": T,
                       "kernel_us_per_env_step": 1e3 * rp["step_ms"] / (reps * T),
                       "hbm_bytes_written_per_env_step": D + 28,
                       "note": "state stays in registers for 8 consecutive env-steps per launch; every step's obs, "
                               "mask, agent and done are stored into slice t of [T,B,...] rollout tensors"}
            del ro
            torch.cuda.empty_cache()

    # end to end through the host-buffer C-ABI entry
    e2e = None
    if args.e2e_steps > 0:
        e2e = run_e2e(args, env, dev, rank, world)
    env.close()
    del env
    torch.cuda.empty_cache()

    # the other reset mode, same workload and timing rules, so that both figures come from one run
    other = None
    if args.other_reset_steps > 0:
        mode2 = "same_step" if args.reset == "next_step" else "next_step"
        r2, _ = measure_step_loop(cx, N, B, args.indirect, mode2, max(1, min(K, args.other_reset_steps)), max(W, 3),
                                  args.preroll, args.seed, profile_steps=256)
        other = {"reset": mode2, "value": r2["value"], "unit": UNIT, "steps": r2["steps"], "ms_per_step": r2["ms_per_step"],
                 "us_per_launch": r2["us_per_launch"], "kernel_us": r2["roofline"]["kernel_us"],
                 "frac": r2["roofline"]["frac"]}

    # BASELINE configs 3 and 5 by the same rules, in the same run (one GPU: config 3; every N: config 5's shard)
    configs = {}
    if not args.no_configs:
        if world == 1:
            K3 = max(7, min(K, 10))         # >= 448 launches of 2^24 8-player envs (~1.07 ms each)
            c3, _ = measure_step_loop(cx, 8, 1 << 24, False, args.reset, K3, 2, args.preroll, args.seed + 3,
                                      profile_steps=64)
            c3["config"] = workload_config(8, 1 << 24, False, args.reset, 1, args)
            c3["baseline_config"] = "configs[2]: 8-player SkyJo, 16M envs on 1 B200"
            configs["c3"] = c3
        K5 = max(7, min(K, 20))
        c5, _ = measure_step_loop(cx, 4, 1 << 23, False, args.reset, K5, 2, args.preroll, args.seed + 5,
                                  profile_steps=128)
        c5["config"] = workload_config(4, 1 << 23, False, args.reset, world, args)
        c5["baseline_config"] = ("configs[4]: 64M envs sharded across 8 B200 with the statistics all-reduce -- 2^23 "
                                 f"envs per GPU, here {world} GPU(s) = {world << 23} envs")
        c5["n_gpus"] = world
        configs["c5"] = c5

    # BASELINE config 4: the action-mask MLP policy consuming obs / mask in place
    policy_rollout = None
    if args.policy_steps > 0 and rank == 0:
        policy_rollout = run_policy_rollout(args, dev)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is timed at N = 1 only
            cpu = cpu_baseline_record(N, args.indirect, args.cpu_budget, with_python=not args.no_python_reference)
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "int8", "data": "synthetic",
            "config": workload_config(N, B, args.indirect, args.reset, world, args),
            "clocks": main["clocks"], "e2e": e2e, "gpu_launches": main["gpu_launches"],
            "graph_replays": main["graph_replays"], "roofline": main["roofline"],
            "cpu_baseline": cpu, "rollout": rollout, "env_steps_per_step": L, "us_per_launch": main["us_per_launch"],
            "timed_region_s": main["timed_region_s"], "counted_env_steps": main["counted_env_steps"],
            "counted_frac": main["counted_frac"], "other_reset_mode": other, "policy_rollout": policy_rollout,
            "configs": configs, "host_cores": os.cpu_count(),
            "stats_allreduce": ("skyjo_stats_allreduce_async: device-side reduction in stream, "
                                + ("in-library ncclAllReduce" if world > 1 else "single-rank copy")
                                + " on a side stream, once per iteration") if (world == 1 or cx.comm is not None)
                               else "torch.distributed.all_reduce",
            "episode_stats": main["episode_stats"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_policy_rollout(args, dev):
    """a torch MLP (TorchActionMaskModel restated, skyjo_rl_b200/policy.py) reads the env's obs /
    mask tensors in place, the fused masked-softmax-sample kernel writes uint8 actions, skyjo_step consumes them.
    One GPU; the GEMMs are ATen's (library code), so this is a consumer-side figure, not a kernel claim."""
    import torch

    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.policy import ActionMaskPolicy

    N, T = args.players, args.policy_steps
    B = min(args.envs, args.policy_envs)
    env = BatchedSkyjoEnv(num_envs=B, num_players=N, observe_other_player_indirect=args.indirect, device=dev,
                          seed=args.seed + 2)
    env.reset()
    torch.manual_seed(0)
    policy = ActionMaskPolicy(env.obs_len).to(dev)
    ptr = (env.observations.data_ptr(), env.action_mask.data_ptr())
    obs = {"observations": env.observations, "action_mask": env.action_mask}

    def steps(n, autocast):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            for _ in range(n):
                act, _ = env.sample_actions(policy(obs).float())
                env.step(act)

    def timed(autocast):
        steps(8, autocast)
        env.clear_stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        ev0.record()
        steps(T, autocast)
        ev1.record()
        torch.cuda.synchronize(dev)
        ms = ev0.elapsed_time(ev1)
        st = env.stats()
        assert st["illegal"] == 0
        return st["steps"] / (ms * 1e-3), ms / T

    value, ms_per_step = timed(False)          # fp32, the reference's precision (RLlib TorchFC)
    value_bf16, ms_bf16 = timed(True)          # the same policy under bf16 autocast (library tensor-core GEMMs)
    out = {"value": value, "unit": UNIT, "envs": B, "steps": T, "ms_per_step": ms_per_step,
           "policy": "ActionMaskPolicy 2x256 tanh + value branch, fp32 (ATen GEMMs), fused masked-softmax-sample kernel",
           "bf16_autocast": {"value": value_bf16, "ms_per_step": ms_bf16},
           "zero_copy": True, "reset": "same_step"}
    # the same loop with the policy side as ONE sm_100a kernel (csrc/skyjo_policy.cu: tcgen05.mma, weights in shared
    # memory, activations in tensor memory, tanh / masked softmax / Philox sample in the epilogues)
    if env.obs_len <= 96:
        from skyjo_rl_b200.policy import FusedPolicy
        fused = FusedPolicy(policy, env, with_value=False)
        acts = torch.empty(B, dtype=torch.uint8, device=dev)
        logp = torch.empty(B, dtype=torch.float32, device=dev)

        def fsteps(n):
            for t in range(n):
                fused.sample(t, acts, logp)
                env.step(acts)

        fsteps(8)
        Tf = max(T, 200)
        env.clear_stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        ev0.record()
        fsteps(Tf)
        ev1.record()
        torch.cuda.synchronize(dev)
        ms = ev0.elapsed_time(ev1)
        st = env.stats()
        assert st["illegal"] == 0
        # the policy kernel alone, back to back on the same observations
        ev0.record()
        for t in range(50):
            fused.sample(t, acts, logp)
        ev1.record()
        torch.cuda.synchronize(dev)
        k_us = 1e3 * ev0.elapsed_time(ev1) / 50
        flop_issued = 2.0 * B * ((96 + 256) * 256 + 256 * 32)       # the MMAs the kernel issues (K padded to 96, N to 32)
        flop_useful = 2.0 * B * (env.obs_len * 256 + 256 * 256 + 256 * 26)
        peak_tf = None
        try:
            peak_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:  # noqa: BLE001
            peak_tf = 1590.0
        out["fused_kernel"] = {
            "value": st["steps"] / (ms * 1e-3), "unit": UNIT, "steps": Tf, "ms_per_step": ms / Tf,
            "kernel": "skyjo::policy_kernel (tcgen05.mma kind::f16, A from tensor memory) + skyjo::step_kernel",
            "policy_kernel_us": k_us, "tflops_issued": flop_issued / (k_us * 1e-6) / 1e12,
            "tflops_useful": flop_useful / (k_us * 1e-6) / 1e12, "peak_bf16_tflops": peak_tf,
            "tensor_frac_issued": flop_issued / (k_us * 1e-6) / 1e12 / peak_tf,
            "sfu_bound_us": B * 512 / (148 * 8 * 1.965e9) * 1e6,
            "note": "bf16 operands, fp32 accumulation; 512 tanh per env on the SFUs (MUFU.TANH: 8 per clock per SM) bound the "
                    "kernel before the tensor cores do (sfu_bound_us at 1965 MHz)"}
    env.check()
    assert ptr == (env.observations.data_ptr(), env.action_mask.data_ptr()), "obs / mask must be consumed in place"
    env.close()
    return out


def run_e2e(args, env, dev, rank, world):
    import torch
    import torch.distributed as dist

    from skyjo_rl_b200 import BatchedSkyjoEnv

    N, B, T = args.players, args.envs, args.e2e_steps
    D = env.obs_len
    kw = dict(num_envs=B, num_players=N, observe_other_player_indirect=args.indirect, device=dev,
              seed=args.seed + 1, first_global_env_id=rank * B, auto_reset=args.reset)
    Tw = 3
    # record a legal action sequence on the device (untimed), then replay it from pinned host memory
    rec = BatchedSkyjoEnv(**kw)
    rec.reset()
    acts = torch.empty((T + Tw, B), dtype=torch.uint8).pin_memory()
    for t in range(T + Tw):
        a = torch.multinomial(rec.action_mask.float(), 1).squeeze(1).to(torch.uint8)
        rec.step(a)
        acts[t].copy_(a)
    torch.cuda.synchronize(dev)
    rec.close()
    del rec
    e = BatchedSkyjoEnv(**kw)
    e.reset()
    obs_h = torch.empty((B, D), dtype=torch.int8).pin_memory()
    mask_h = torch.empty((B, 26), dtype=torch.int8).pin_memory()
    agent_h = torch.empty(B, dtype=torch.int8).pin_memory()
    done_h = torch.empty(B, dtype=torch.uint8).pin_memory()
    rew_h = torch.empty((B, N), dtype=torch.float64).pin_memory()
    for t in range(Tw):
        e.step_host(acts[t], obs_h, mask_h, agent_h, done_h, rew_h)
    e.clear_stats()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for t in range(Tw, Tw + T):
        e.step_host(acts[t], obs_h, mask_h, agent_h, done_h, rew_h)   # synchronises every step
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    st = e.stats(all_reduce=world > 1)       # the timed steps only
    e.check()
    assert st["illegal"] == 0, "replayed actions must be legal"
    # bytes that cross the link per step (csrc/skyjo_hostio.cuh), as counted by the library from the copies it
    # queues: the observation row and one packed mask + agent + done word per env, an 8-byte count; plus one
    # {env, N rewards} entry (written by the pack kernel straight into host-mapped memory) per episode that ended
    ended_per_step = (st["episodes"] + st["truncated"]) / float(T * world)
    d2h = e.host_wire_bytes + int(ended_per_step * (8 + 8 * N))
    out = {"value": st["steps"] / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": B,
           "d2h_bytes_per_step": d2h, "steps": T, "ms_per_call": 1e3 * float(dt.item()) / T,
           "host_buffer_bytes_filled_per_step": B * (D + 26 + 1 + 1 + 8 * N),
           "wire_bytes_per_env": round(e.host_wire_bytes / float(B), 2), "compact_ranges_of_8": e.host_wire_share,
           "api": "skyjo_step_host (C ABI) via BatchedSkyjoEnv.step_host, pinned host buffers: obs int8[B,D], "
                  "mask int8[B,26], agent, done, reward f64[B,N] all filled every step; one call = one env-step of "
                  "every env (wall clock around the calls, max over ranks)"}
    e.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60, help="timed iterations of --launches-per-step env-step launches")
    ap.add_argument("--warmup", type=int, default=5, help="untimed iterations (after --preroll single steps)")
    ap.add_argument("--launches-per-step", type=int, default=64,
                    help="env-step launches per iteration (one statistics all-reduce closes each iteration)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--players", type=int, default=4)
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--indirect", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--preroll", type=int, default=640)
    ap.add_argument("--reset", default="next_step", choices=["same_step", "next_step"],
                    help="auto-reset mode (SKYJO_RESET_*): next_step = phase-locked, reset slots are not counted")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--rollout-steps", type=int, default=64, help="rollout length T of the multi-step path (0 = skip)")
    ap.add_argument("--other-reset-steps", type=int, default=20,
                    help="also time up to this many iterations in the other auto-reset mode (0 = skip)")
    ap.add_argument("--policy-steps", type=int, default=24, help="steps of the torch-policy rollout, config 4 (0 = skip)")
    ap.add_argument("--policy-envs", type=int, default=1 << 18)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="CPU-seconds of oracle work (baseline sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config 3 / 5 sub-records")
    ap.add_argument("--global-envs", type=int, default=0,
                    help="strong scaling (SURVEY 8d C5): total envs fixed, split evenly over the GPUs (overrides --envs)")
    args = ap.parse_args()
    args.scaling = "weak"
    if args.global_envs > 0:
        args.envs = args.global_envs // max(int(os.environ.get("WORLD_SIZE", 1)), 1)
        args.scaling = "strong"
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    # libraries (NCCL's version banner) write to fd 1: keep stdout for the one JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
