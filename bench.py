#!/usr/bin/env python3
"""Benchmark of the hot path: batched KS environment stepping fused with policy inference.

Metric (BASELINE.json): env-steps/s, KS N=256, 8192 batched envs per GPU, oversampling 30,
fp64 (the reference's precision), actor = the reference's shipped KS200 network (1->6->1).
One "step" = one {actor forward -> prepare_action -> 30 CNAB2 substeps -> reward -> featurize -> done} pass
over all environments of the rank = THREE kernel launches on one stream (actuation, KS core, observation).

The line also carries a `train` record (BASELINE config 5): the KS training loop -- policy with exploration noise,
replay push, update_loops x {sample, critic, actor} as one CUDA graph with the gradient exchange over NVLink peer
memory inside the kernels, env step, reward push -- at update_loops 1 and 20, on every N.

  python bench.py --gpus N --steps K --warmup W            our arm (under torchrun for N>1)
  python bench.py --impl reference ...                     CPU arm: oracle restatement on all host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is produced.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
PKG = "distributedconvrl-pde-control_b200"

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=8192, help="environments per GPU (weak scaling)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--oversampling", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-loop record (BASELINE config 5)")
    ap.add_argument("--train-batch", type=int, default=4096, help="DDPG batch (replay columns) per GPU and update")
    ap.add_argument("--train-steps", type=int, default=30, help="timed loop steps per update_loops setting")
    ap.add_argument("--e2e-driver", default="native", choices=["native", "python"],
                    help="host threads of the e2e leg: native std::threads calling the C ABI (libpdeb200_host.so) or Python threads")
    ap.add_argument("--e2e-calls", default="device-agent", choices=["device-agent", "fused", "split"],
                    help="device-agent: the drop-in as deployed (policy and trajectory on the device): per step host noise in "
                         "[H2D], the packed result block (--e2e-result) out in one copy [D2H], one C call, one sync; fused: the same call "
                         "with the action additionally round-tripping through the host; split: policy_act -> get(ACTION_IN) -> "
                         "step_host (three syncs, four copies; round 1's sequence)")
    ap.add_argument("--e2e-result", default="reward+done", choices=["reward+done", "full"],
                    help="device-agent flow: what the packed per-step D2H copy carries.  reward+done: what host code reads every "
                         "step when policy and trajectory are on the device (the hook reads env.reward, the run loop env.done; "
                         "pdeb200_result_select(ctx, 0)); full: [reward | done | state].  The other one is measured too and "
                         "reported as e2e.other_result")
    ap.add_argument("--e2e-noise", default="prefetch", choices=["prefetch", "inline"],
                    help="device-agent flow: prefetch = step i+1's host noise is handed to pdeb200_noise_prefetch before the call for "
                         "step i (its upload overlaps step i's kernels); inline = uploaded at the head of its own call")
    ap.add_argument("--e2e-shards", type=int, default=4,
                    help="the e2e leg drives the batch as this many env shards (own context + stream + host thread each) "
                         "so that one shard's PCIe copies overlap another shard's kernels")
    return ap.parse_args()


def config_dict(args, n_gpus):
    return {"workload": "KS 1D N=256 (Lx=213.33, dt=0.1, oversampling=%d), %d envs/GPU, 64 sensors/actuators, window 1, "
                        "fused policy inference (KS200 actor 1-6-1) + env step" % (args.oversampling, args.envs),
            "envs_per_gpu": args.envs, "global_envs": args.envs * n_gpus, "nx": 256, "oversampling": args.oversampling,
            "parallelism": "env-sharded x%d (no data-path collective)" % n_gpus,
            "l2": "flushed between timed iterations (256 MiB write, then read back so the evicted-to lines are clean); state 16 MiB/GPU < 126 MB L2"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
class CpuOracle:
    def __init__(self, oversampling):
        from oracle import ks_oracle as K
        so = ROOT / "oracle" / "_ref" / "libks_oracle.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(ROOT / "oracle"), "-s"], check=True)
        self.lib = C.CDLL(str(so))
        self.K = K
        self.cfg = K.ks256_config(1)
        self.cfg.oversampling = oversampling
        env = K.KSEnv(self.cfg)
        self.g_sens = np.ascontiguousarray(env.g_sens)
        self.g_act = np.ascontiguousarray(env.g_act)
        self.a2s = np.ascontiguousarray(np.asarray(self.cfg.actuators_to_sensors) - 1, dtype=np.int32)
        self.cores = len(os.sched_getaffinity(0))

    def run(self, envs_per_core, n_steps, seed=0):
        """Advance cores*envs_per_core environments n_steps env steps; one thread per core, each
        stepping its environments one at a time like the (single-threaded) reference.  Returns seconds."""
        cfg, K = self.cfg, self.K
        rng = np.random.default_rng(seed)
        nx, n_a, n_s = cfg.nx, cfg.n_actuators, cfg.n_sensors
        slices = []
        for c in range(self.cores):
            y = np.stack([K.generate_random_init(cfg, rng.uniform(-1, 1, 8)) for _ in range(envs_per_core)])
            ap = np.zeros((envs_per_core, n_a))
            act = rng.uniform(-0.2, 0.2, (n_steps, envs_per_core, n_a))
            st = np.zeros((envs_per_core, n_a, 1))
            rw = np.zeros((envs_per_core, n_a))
            slices.append((y, ap, act, st, rw))
        d = lambda a: a.ctypes.data_as(C.c_void_p)

        def work(s):
            y, ap, act, st, rw = s
            self.lib.ks_oracle_env_steps(
                C.c_int(nx), C.c_double(cfg.Lx), C.c_double(cfg.dt), C.c_int(cfg.oversampling), C.c_double(cfg.mu),
                C.c_int(envs_per_core), C.c_int(n_steps), C.c_int(n_s), C.c_int(n_a), C.c_int(1),
                C.c_double(cfg.agent_power), C.c_double(cfg.max_value), C.c_double(cfg.action_punish),
                C.c_double(cfg.delta_action_punish), d(self.g_sens), d(self.g_act), d(self.a2s), d(y), d(ap), d(act),
                d(st), d(rw), C.c_int(1))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=self.cores) as ex:
            list(ex.map(work, slices))
        return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc = CpuOracle(args.oversampling)
    # one bench "step" of this arm = a bounded sample: 64 envs per core advanced 4 env steps (large enough that the
    # thread start-up does not count against the CPU; ~0.15 s per step, a 200-step run stays under a minute)
    envs_per_core, inner = 64, 4
    for _ in range(args.warmup):
        orc.run(envs_per_core, inner)
    t = 0.0
    for i in range(args.steps):
        t += orc.run(envs_per_core, inner, seed=i + 1)
    n_env_steps = args.steps * envs_per_core * orc.cores * inner
    value = n_env_steps / t
    sample = "%d envs (%d per core) x %d env steps per bench step" % (envs_per_core * orc.cores, envs_per_core, inner)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": orc.cores, "kind": "port", "sample": sample,
                             "note": "C restatement of the reference algorithm (oracle/ks_oracle.c, literal 123 FFTs/"
                                     "env-step, own mixed-radix DFT), not Julia+FFTW: Julia is not installed"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions (NVML from a thread, every ~5 ms;
    `nvidia-smi -lms` as the fallback)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.sm, self.reasons, self.power, self.stop_flag, self.mx = [], set(), [], False, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if self.index < len(ids) and ids[self.index].strip().isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.004)

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm),
                    "power_w_max": max(self.power) if self.power else None, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def bind_to_gpu_cpus(index):
    """Bind this rank to the CPUs NVML reports as local to its GPU (intersected with the cpuset we are allowed):
    the e2e leg's pinned buffers and driver threads then sit on the GPU's own NUMA node / PCIe root."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(ids) and ids[index].strip().isdigit():
                phys = int(ids[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(local & allowed)
        if use:
            os.sched_setaffinity(0, use)
            return {"gpu_local_cpus": len(local), "bound_to": len(use)}
        return {"gpu_local_cpus": len(local), "bound_to": 0, "note": "no GPU-local CPU in the allowed cpuset"}
    except Exception as e:                                   # noqa: BLE001 -- best effort, never fatal
        return {"error": type(e).__name__}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def train_record(args, pkg, torch, dist, rank, world, local):
    """BASELINE config 5: KS training loop at `world` GPUs.  One loop step = policy (actor + device Philox noise) ->
    PreAct push of (s, a) for every column -> update_loops x {sample, critic grads + exchange + ADAM, actor grads +
    exchange + ADAM + Polyak} (ONE CUDA graph launch; the exchange is the library's peer-memory allreduce inside the
    gradient kernels) -> env step -> PostAct push of (r, terminal) -> masked reset of diverged environments; the stage
    order of RLCore's run() (scripts/Fluid/setup/FluidSetup.jl:455-519, src/PDEagent.jl:342-361).  Device events on the
    context's stream, max over ranks; weights must be bit-identical on every rank afterwards."""
    A, L, par = pkg.agent, pkg.lib, pkg.parallel
    B = args.envs
    setup = pkg.setups.KSSetup.ks256(oversampling=args.oversampling)
    out = {"workload": "KS N=256 training loop (BASELINE config 5): %d envs/GPU, DDPG batch %d replay columns/GPU per update, "
                       "actor 1-6-1 / critic 2-140-1 (KSSetup.jl sizes), act_noise 1.2, literal quirk-Q1 loss, %s"
                       % (B, args.train_batch, args.dtype),
           "unit": UNIT, "n_gpus": world, "steps": args.train_steps}
    comm = par.Comm(dist if world > 1 else None)
    for loops in (1, 20):
        rng = np.random.default_rng(100 + rank)
        env = setup.make_env(n_envs=B, dtype=args.dtype, device=local, y0=setup.generate_random_init(rng, B))
        stream = torch.cuda.current_stream()
        L.check(env._lib.pdeb200_set_stream(env._ctx, C.c_void_p(stream.cuda_stream)), env._ctx)
        wrng = np.random.default_rng(7)                                 # identical initial weights on every rank
        pol = A.create_agent(env, rng=wrng, nna_scale=0.6, nna_scale_critic=7.0, drop_middle_layer=True,
                             batch_size=args.train_batch, start_steps=2, update_after=2, update_freq=1, update_loops=loops,
                             act_noise=1.2, trajectory_length=B * env.n_cols * 8, seed=rank,
                             comm=comm if world > 1 else None)
        traj = pol.trajectory
        env.reset()
        traj.pre_episode()

        def step():
            pol(env, learning=True)
            traj.pre_act()
            pol.maybe_update()
            env.step_device()
            traj.post_act()
            env.reset_diverged(sync=False)

        for _ in range(8):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0, u0 = env.launch_count, pol.n_updates
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.train_steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_local = e0.elapsed_time(e1)
        per_rank = [ms_local]
        w = np.concatenate([n.sync_from_device().flat() for n in
                            (pol.behavior_critic, pol.behavior_actor, pol.target_critic, pol.target_actor)])
        identical = True
        if world > 1:
            t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
            allms = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allms, t)
            per_rank = [float(x.item()) for x in allms]
            wt = torch.from_numpy(w).cuda()
            allw = [torch.zeros_like(wt) for _ in range(world)]
            dist.all_gather(allw, wt)
            identical = all(bool(torch.equal(x, allw[0])) for x in allw[1:])
        ms = max(per_rank)
        rec = {"value": B * world * args.train_steps / (ms * 1e-3), "ms_per_loop_step": ms / args.train_steps,
               "per_rank_ms_per_loop_step": [x / args.train_steps for x in per_rank],
               "updates_per_step": (pol.n_updates - u0) / args.train_steps,
               "gpu_launches_per_step": (env.launch_count - l0) / args.train_steps,
               "weights_finite": bool(np.all(np.isfinite(w))), "weights_identical_across_ranks": identical,
               "losses": pol.losses}
        out["update_loops_%d" % loops] = rec
        out["transport"] = {L.COMM_NONE: "none (single GPU)", L.COMM_NCCL: "nccl allreduce between phases",
                            L.COMM_PEER: "NVLink peer-memory exchange inside the gradient kernels"}[pol.transport]
        env.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_cpus(local)       # pinned host buffers are then first-touched on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module(PKG)
    agent = importlib.import_module(PKG + ".agent")
    L = pkg.lib

    B = args.envs
    setup = pkg.setups.KSSetup.ks256(oversampling=args.oversampling)
    rng = np.random.default_rng(1000 + rank)
    y0 = setup.generate_random_init(rng, B)
    env = setup.make_env(n_envs=B, dtype=args.dtype, device=local, y0=y0)
    gpath = ROOT / "tests" / "golden" / "ks200_hook.npz"
    g = np.load(gpath)
    chain = agent.Chain(agent.Dense(g["best_W1"], g["best_b1"], "relu"), agent.Dense(g["best_W2"], g["best_b2"], "tanh"))
    agent.CustomNeuralNetworkApproximator(env, L.NET_BEHAVIOR_ACTOR, chain)
    # a real (non-default) torch stream: the default stream's handle is NULL, which the C ABI reads as
    # "use the context's own stream"; kernels, flushes and events must share one stream to be timed
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    L.check(env._lib.pdeb200_set_stream(env._ctx, C.c_void_p(stream.cuda_stream)), env._ctx)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush_sink = torch.zeros(1, dtype=torch.int64, device="cuda")

    def flush_l2():
        # write a buffer larger than L2 (evicts everything), then read it back so that L2 is left holding CLEAN
        # lines: a timed step should start cold, not pay for writing someone else's dirty 126 MB back to HBM
        flush.zero_()
        flush_sink.copy_(flush.view(torch.int64).sum().reshape(1))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: one rollout(1) call = three launches per step ------------------
    for _ in range(max(args.warmup, 3)):
        env.rollout(1)
    torch.cuda.synchronize()
    env.reset()
    fma_peak = env.measure_fma_peak()            # measured CUDA-core peak of THIS GPU for the compute dtype (TFLOP/s)
    launches0 = env.launch_count
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    for i in range(args.steps):
        flush_l2()                         # L2 flush, outside the event-timed region
        ev0[i].record(stream)
        env.rollout(1)
        ev1[i].record(stream)
    barrier()
    launches = env.launch_count - launches0
    kern_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    total_ms = float(sum(kern_ms))
    # dominant kernel (the KS core kernel) alone: CUDA events recorded around it inside the library, same stream
    L.check(env._lib.pdeb200_enable_step_timing(env._ctx, 1), env._ctx)
    core_ms, phases = [], []
    for i in range(min(args.steps, 50)):
        flush_l2()
        env.rollout(1)
        ms = C.c_float()
        L.check(env._lib.pdeb200_last_core_ms(env._ctx, C.byref(ms)), env._ctx)
        core_ms.append(ms.value)
        ph = (C.c_float * 3)()
        L.check(env._lib.pdeb200_last_phase_ms(env._ctx, ph), env._ctx)
        phases.append([ph[0], ph[1], ph[2]])
    phases = [float(x) for x in np.mean(np.asarray(phases), axis=0)]
    L.check(env._lib.pdeb200_enable_step_timing(env._ctx, 0), env._ctx)
    core_ms = float(np.mean(core_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    per_rank_ms = [total_ms / args.steps]
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x.item()) / args.steps for x in allt]          # names the straggler when efficiency < 1
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = B * world * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the public API with HOST buffers -----------------------------------
    # Per step, as the drop-in closures do it: policy(env) returns the action on the host
    # (PDEagent.jl:198), env(action) takes it from the host, reward/state/done come back for the
    # agent and the hook.
    esz = 8 if args.dtype == "f64" else 4
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    n_sh = max(1, args.e2e_shards)
    while B % n_sh:
        n_sh -= 1
    Bs = B // n_sh
    env.close()                                   # the e2e shards replace the device-timed context

    class Shard:
        def __init__(self, k):
            self.env = setup.make_env(n_envs=Bs, dtype=args.dtype, device=local, y0=y0[k * Bs:(k + 1) * Bs])
            agent.CustomNeuralNetworkApproximator(self.env, L.NET_BEHAVIOR_ACTOR, chain.copy())
            self.n_act = Bs * self.env.n_actuators
            self.h_act = torch.empty(self.n_act, dtype=tdt).pin_memory()
            self.h_rew = torch.empty(self.n_act, dtype=tdt).pin_memory()
            self.h_state = torch.empty(self.n_act * self.env.ns, dtype=tdt).pin_memory()
            self.h_done = torch.empty(Bs, dtype=torch.uint8).pin_memory()
            tot = C.c_size_t()
            L.check(self.env._lib.pdeb200_result_layout(self.env._ctx, None, None, None, C.byref(tot)), self.env._ctx)
            self.packed_bytes = tot.value
            self.h_packed = torch.empty(tot.value, dtype=torch.uint8).pin_memory()      # [reward | done | state], one D2H
            # exploration noise drawn on the host like the reference's randn(policy.rng, ...) (PDEagent.jl:201): always float64
            self.h_noise = torch.from_numpy(np.random.default_rng(7 + k).standard_normal(self.n_act)).pin_memory()
            self.primed = False

        def step(self):
            lib, ctx = self.env._lib, self.env._ctx
            if args.e2e_calls == "fused":
                L.check(lib.pdeb200_act_step_host(ctx, None, 0.0, 1.0, C.c_void_p(self.h_act.data_ptr()), None,
                                                  C.c_void_p(self.h_packed.data_ptr()), None, None, None), ctx)
                return
            if args.e2e_calls == "device-agent":
                if args.e2e_noise == "prefetch":
                    if not self.primed:
                        L.check(lib.pdeb200_noise_prefetch(ctx, C.c_void_p(self.h_noise.data_ptr())), ctx)
                        self.primed = True
                    L.check(lib.pdeb200_noise_prefetch(ctx, C.c_void_p(self.h_noise.data_ptr())), ctx)
                    L.check(lib.pdeb200_act_step_host(ctx, None, 0.05, 1.0, None, None, C.c_void_p(self.h_packed.data_ptr()), None, None, None), ctx)
                    return
                L.check(lib.pdeb200_act_step_host(ctx, C.c_void_p(self.h_noise.data_ptr()), 0.05, 1.0, None, None,
                                                  C.c_void_p(self.h_packed.data_ptr()), None, None, None), ctx)
                return
            L.check(lib.pdeb200_policy_act(ctx, None, 0.0, 1.0), ctx)
            L.check(lib.pdeb200_get(ctx, L.ARR_ACTION_IN, C.c_void_p(self.h_act.data_ptr()), self.n_act * esz), ctx)
            L.check(lib.pdeb200_step_host(ctx, C.c_void_p(self.h_act.data_ptr()), None, C.c_void_p(self.h_rew.data_ptr()),
                                          C.c_void_p(self.h_state.data_ptr()), C.c_void_p(self.h_done.data_ptr())), ctx)

    shards = [Shard(k) for k in range(n_sh)]

    host_lib = None
    if args.e2e_driver == "native":
        hp = ROOT / PKG / "libpdeb200_host.so"
        if not hp.exists():
            raise SystemExit("bench.py: %s missing (python __graft_entry__.py build)" % hp)
        host_lib = C.CDLL(str(hp))
        host_lib.pdeb200_host_drive.restype = C.c_int32
        host_lib.pdeb200_host_drive2.restype = C.c_int32
        VP = C.c_void_p * n_sh
        packs = VP(*[sh.h_packed.data_ptr() for sh in shards])
        noises = VP(*[sh.h_noise.data_ptr() for sh in shards])
        ctxs = VP(*[sh.env._ctx.value for sh in shards])
        acts = VP(*[sh.h_act.data_ptr() for sh in shards])
        rews = VP(*[sh.h_rew.data_ptr() for sh in shards])
        sts = VP(*[sh.h_state.data_ptr() for sh in shards])
        dns = VP(*[sh.h_done.data_ptr() for sh in shards])
        nbytes = (C.c_size_t * n_sh)(*[sh.n_act * esz for sh in shards])

    def drive(n):
        if host_lib is not None:
            # one std::thread per shard, each running policy_act -> get(ACTION_IN) -> step_host through the C ABI
            secs = C.c_double()
            if args.e2e_calls == "fused":
                rc = host_lib.pdeb200_host_drive2(C.c_int32(n_sh), ctxs, C.c_int32(n), acts, packs, C.c_double(1.0), C.byref(secs),
                                                  None, C.c_double(0.0))
            elif args.e2e_calls == "device-agent":
                rc = host_lib.pdeb200_host_drive2(C.c_int32(n_sh), ctxs, C.c_int32(n), None, packs, C.c_double(1.0), C.byref(secs),
                                                  noises, C.c_double(-0.05 if args.e2e_noise == "prefetch" else 0.05))
            else:
                rc = host_lib.pdeb200_host_drive(C.c_int32(n_sh), ctxs, C.c_int32(n), acts, nbytes, rews, sts, dns, C.c_double(1.0),
                                                 C.byref(secs))
            if rc:
                raise SystemExit("bench.py: e2e driver failed with %d" % rc)
            return secs.value
        # ctypes releases the GIL inside each C-ABI call, so the shard threads really overlap
        def loop(sh):
            for _ in range(n):
                sh.step()
        t_ = time.perf_counter()
        if n_sh == 1:
            loop(shards[0])
        else:
            ths = [threading.Thread(target=loop, args=(sh,)) for sh in shards]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        return time.perf_counter() - t_

    def timed_e2e(with_state):
        """-> (env-steps/s over all ranks, packed D2H bytes per step and GPU) with the packed copy carrying the whole
        [reward | done | state] block or its [reward | done] prefix"""
        tot = C.c_size_t()
        for sh in shards:
            L.check(sh.env._lib.pdeb200_result_select(sh.env._ctx, 1 if with_state else 0), sh.env._ctx)
            L.check(sh.env._lib.pdeb200_result_layout(sh.env._ctx, None, None, None, C.byref(tot)), sh.env._ctx)
            sh.packed_bytes = tot.value
        drive(3)
        barrier()
        t0_ = time.perf_counter()
        drive(args.steps)
        torch.cuda.synchronize()
        t_ = torch.tensor([time.perf_counter() - t0_], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return B * world * args.steps / float(t_.item()), sum(sh.packed_bytes for sh in shards)

    primary_full = args.e2e_result == "full" or args.e2e_calls != "device-agent"
    other_result = None
    if args.e2e_calls == "device-agent":
        ov, ob = timed_e2e(not primary_full)
        other_result = {"result": "[reward | done | state]" if not primary_full else "[reward | done]", "value": ov, "d2h_bytes_per_step": ob}
    e2e_launches0 = sum(sh.env.launch_count for sh in shards)
    e2e_value, packed_total = timed_e2e(primary_full)
    clk = clocks.stop()
    n_act = B * shards[0].env.n_actuators
    ns_rows = shards[0].env.ns
    if args.e2e_calls == "device-agent":
        h2d = n_act * 8                                           # exploration noise (float64 like the reference's randn)
        d2h = packed_total                                         # packed blocks (256-byte aligned parts)
    else:
        h2d = n_act * esz
        d2h = n_act * esz + n_act * esz + n_act * ns_rows * esz + B
    e2e_launches = sum(sh.env.launch_count for sh in shards) - e2e_launches0
    env = shards[0].env                             # step_cost below is per environment

    # ---- roofline of the dominant (only) kernel -------------------------------------------------
    bytes_env, flops_env = env.step_cost()
    avg_ms = float(np.mean(kern_ms))
    peak, peak_src = measured_peak()
    # the core kernel's own algorithmic bytes: y in + y out + p in + sensor dots out; the step's: bytes_env
    esz_ = 8 if args.dtype == "f64" else 4
    core_bytes_env = (3 * 256 + 64) * esz_
    achieved = core_bytes_env * B / (core_ms * 1e-3) / 1e9
    core_kernel = (env._lib.pdeb200_last_core_kernel(env._ctx) or b"").decode() or "ks_step_kernel"
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("%s|%s|S%d|%d envs" % (core_kernel.split("<")[0], args.dtype, args.oversampling, B))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": "profiles/traffic.json (one ncu --set full capture of this kernel and config; "
                "not re-measured in this run)" if traffic is not None else None, "peak_source": peak_src, "kernel": core_kernel,
                "algorithmic_bytes_per_launch": core_bytes_env * B, "kernel_ms": core_ms,
                "kernel_share_of_step": core_ms / avg_ms,
                "phase_ms": {"actuate": phases[0], "core": phases[1], "observe": phases[2]},
                "step": {"algorithmic_bytes_per_env_step": bytes_env, "ms": avg_ms,
                         "achieved_gbs": bytes_env * B / (avg_ms * 1e-3) / 1e9, "frac": bytes_env * B / (avg_ms * 1e-3) / 1e9 / peak},
                "note": "at oversampling=%d the step is FP64-pipe/shared-memory bound (arithmetic intensity ~%d flop/B against a ridge of "
                        "~5.6), so the HBM fraction is small by construction; see fp_pipe" % (args.oversampling, round(flops_env / bytes_env)),
                "binding_roof": "fp_pipe (%s CUDA-core FMA)" % args.dtype,
                "fp_pipe": {"algorithmic_flops_per_env_step": flops_env,
                            "achieved_tflops": flops_env * B / (core_ms * 1e-3) / 1e12,
                            "measured_peak_tflops": fma_peak,
                            "peak_source": "pdeb200_measure_fma_peak: FMA micro-kernel on this GPU, same run",
                            "frac": flops_env * B / (core_ms * 1e-3) / 1e12 / fma_peak if fma_peak else None,
                            "frac_of_whole_step": flops_env * B / (avg_ms * 1e-3) / 1e12 / fma_peak if fma_peak else None,
                            "nominal_peak_tflops": 37.0 if args.dtype == "f64" else 75.0}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "per_rank_ms_per_step": per_rank_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": config_dict(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "shards": n_sh, "launches": int(e2e_launches), "host_threads": args.e2e_driver, "cpu_binding": numa,
                    "result": ("[reward | done | state]" if primary_full else "[reward | done]") if args.e2e_calls == "device-agent" else "reward, state, done",
                    "other_result": other_result,
                    "note": "the shards step continuously, so one shard's actuation / observation kernels and copies run under another "
                            "shard's core kernel; `value` times isolated steps of one 8192-environment context (L2 flushed in between), "
                            "which is why this leg can exceed it when its copies are small",
                    "call_sequence": (("per shard and step: pdeb200_noise_prefetch(next step's host noise) [H2D, overlapping this step's kernels] + " if args.e2e_noise == "prefetch" else "per shard and step: ") +
                                      "ONE call pdeb200_act_step_host = host-drawn exploration noise in [H2D] -> "
                                      "policy(env) on the device -> env(action) -> " + ("[reward | done | state]" if primary_full else "[reward | done] (what host code reads per step: the hook env.reward, the run loop env.done; the observation is consumed by the device policy and pushed to the device trajectory)") + " to the host in one packed copy "
                                      "[D2H]; one synchronisation; pinned host buffers; action and replay stay on the device "
                                      "(DevicePolicyForward / DeviceTrajectory of the Julia shim)") if args.e2e_calls == "device-agent" else
                                     ("per shard and step: pdeb200_act_step_host = policy(env) -> action to the host [D2H] -> env(action) "
                                      "from the host [H2D] -> [reward | done | state] to the host in one packed copy [D2H]; one "
                                      "synchronisation; pinned host buffers") if args.e2e_calls == "fused" else
                                     ("per shard and step: pdeb200_policy_act -> pdeb200_get(ACTION_IN) [D2H] -> "
                                      "pdeb200_step_host [H2D action; D2H reward, state, done], pinned host buffers")},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        orc = CpuOracle(args.oversampling)
        orc.run(4, 1)
        n_steps_cpu = 100                # ~10 s of CPU work on 16 cores (contract: a bounded 10-30 s sample)
        per_core = 128
        secs = orc.run(per_core, n_steps_cpu)
        line["cpu_baseline"] = {"value": per_core * orc.cores * n_steps_cpu / secs, "unit": UNIT, "cores": orc.cores,
                                "kind": "port", "sample": "%d envs x %d env steps (same KS N=256 config), one thread per core"
                                % (per_core * orc.cores, n_steps_cpu)}
    elif rank == 0:
        line["cpu_baseline"] = None
    for sh in shards:
        sh.env.close()
    if not args.no_train:
        line["train"] = train_record(args, pkg, torch, dist, rank, world, local)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
