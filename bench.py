#!/usr/bin/env python3
"""bench.py -- env steps/s of the batched humanoid walk rollout (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs-per-gpu E]

A "step" is one pass of the hot path over one batch: every env of the rank takes one env step
(PD/torque -> RK4 mj_step with collision + PGS -> mocap reward -> termination -> auto reset), i.e.
E env-steps per GPU per step.  Workload at N=1: BASELINE configs[1], "4096-env walk imitation,
random-action rollout, 1xB200" (full 5-term imitation reward, CoM termination, RSI auto-reset);
N>1 is weak scaling (E envs per GPU) with one NCCL all-gather of the (obs, reward, done) record per
step, included in the timed region.

Output: ONE JSON line on rank 0 (see the task contract): value = whole-job env steps/s with state
resident in HBM (CUDA-event time, max over ranks); e2e = same metric through DPVecEnv.step with host
(pinned) action buffers copied H2D and the record copied D2H inside the timed region; roofline =
algorithmic bytes (1196 B per env-step) / measured kernel time vs the measured HBM
peak; cpu_baseline = the float64 oracle port timed on the box's host cores on a bounded sample.

--impl reference times the reference's CPU implementation of the path.  mujoco-py / MuJoCo 2.0 are
not installable here (closed binary, no network), so this is the oracle port (oracle/dm_oracle.c),
one env per host core on all cores -- the one other place that may execute oracle/.
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

ALG_BYTES_PER_ENV_STEP = 1196.0  # 540 B read + 656 B written per env-step
METRIC = "env steps/sec (batched humanoid walk)"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--motion", default="walk")
    ap.add_argument("--reward-mode", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    return ap.parse_args()


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    """Cores this process may actually use: CPU affinity, capped by the cgroup CPU quota when one is set
    (oversubscribing a quota-limited container only adds context switches to the CPU arm)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:            # cgroup v2: "<quota|max> <period>"
            q, per = f.read().split()
        if q != "max":
            n = max(1, min(n, int(-(-int(q) // int(per)))))
    except Exception:
        try:                                                  # cgroup v1
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = max(1, min(n, -(-q // per)))
        except Exception:
            pass
    return n


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML every 10 ms when pynvml is
    importable (nvidia_ml_py), else one nvidia-smi query per 200 ms."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()   # samples: (sm_mhz, max_mhz, [reasons])
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.bits = [(getattr(pynvml, n), k) for n, k in (
                ("nvmlClocksEventReasonHwSlowdown", "hw_slowdown"), ("nvmlClocksEventReasonHwThermalSlowdown", "hw_thermal_slowdown"),
                ("nvmlClocksEventReasonSwThermalSlowdown", "sw_thermal_slowdown"), ("nvmlClocksEventReasonSwPowerCap", "sw_power_cap"))
                         if hasattr(pynvml, n)]
            if not self.bits:
                self.bits = [(getattr(pynvml, n), k) for n, k in (
                    ("nvmlClocksThrottleReasonHwSlowdown", "hw_slowdown"), ("nvmlClocksThrottleReasonHwThermalSlowdown", "hw_thermal_slowdown"),
                    ("nvmlClocksThrottleReasonSwThermalSlowdown", "sw_thermal_slowdown"), ("nvmlClocksThrottleReasonSwPowerCap", "sw_power_cap"))
                             if hasattr(pynvml, n)]
        except Exception:
            self.nv = None

    def _sample_nvml(self):
        nv = self.nv
        sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        self.samples.append((int(sm), int(mx), [k for b, k in self.bits if mask & b]))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 6 and f[0].isdigit():
            self.samples.append((int(f[0]), int(f[1]) if f[1].isdigit() else 0,
                                 [self.NAMES[i] for i in range(4) if f[2 + i].lower().startswith("active")]))

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nv is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(0.01 if self.nv is not None else 0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = sorted(s[0] for s in self.samples)
        mx = max([s[1] for s in self.samples] or [0])
        reasons = sorted({r for s in self.samples for r in s[2]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nv is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------
_W = {}


def _worker_init(motion, reward_mode):
    """Per-process persistent oracle env (one env per host core)."""
    import ctypes as C
    import multiprocessing as mp
    import oracle.pyoracle as po
    from deepmimic_mujoco_b200.model_blob import default_config, pack_model
    from deepmimic_mujoco_b200.refaux import compute_ref_aux
    from deepmimic_mujoco_b200.sim import default_model_tables, load_motions, make_mocap_struct
    ident = mp.current_process()._identity
    wid = ident[0] if ident else 0
    m = pack_model(default_model_tables(), max_con=16, max_efc=40)
    cfg = default_config(reward_mode=reward_mode, auto_reset=1)
    aux = compute_ref_aux([motion]) if reward_mode == 4 else None
    mcs, keep = make_mocap_struct(load_motions([motion]), aux)
    L = po.lib()
    e = po.DmoEnv()
    L.dmo_env_init(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 1234, wid, 0)
    L.dmo_env_reset(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 0)
    L.dmo_rollout(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 300, 99)  # warm-up
    _W.update(L=L, m=m, cfg=cfg, mcs=mcs, keep=keep, e=e, C=C, t=0)


def _worker_run(nsteps):
    C = _W["C"]
    _W["t"] += 1
    t0 = time.perf_counter()
    _W["L"].dmo_rollout(C.byref(_W["m"]), C.byref(_W["cfg"]), C.byref(_W["mcs"]), C.byref(_W["e"]), nsteps, 7 + _W["t"])
    return nsteps, time.perf_counter() - t0


class CpuRollout:
    """The float64 oracle port on `cores` host cores, one persistent env per core."""

    def __init__(self, cores, motion, reward_mode):
        import multiprocessing as mp
        self.cores = cores
        self.pool = None
        if cores == 1:
            _worker_init(motion, reward_mode)
        else:
            self.pool = mp.get_context("fork").Pool(cores, initializer=_worker_init, initargs=(motion, reward_mode))

    def run(self, nsteps):
        res = [_worker_run(nsteps)] if self.pool is None else self.pool.map(_worker_run, [nsteps] * self.cores, chunksize=1)
        total = sum(r[0] for r in res)
        wall = max(r[1] for r in res)
        return total / wall, total, wall

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = 1000  # env-steps per core per bench "step": a bounded sample of the same workload
    cpu = CpuRollout(cores, a.motion, a.reward_mode)
    t_all, n_all, nst = 0.0, 0, 0
    t_begin = time.perf_counter()
    for i in range(a.warmup + a.steps):
        v, n, w = cpu.run(per_step)
        if i >= a.warmup:
            t_all += w; n_all += n; nst += 1
        if time.perf_counter() - t_begin > 150:
            break
    cpu.close()
    value = n_all / t_all
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": nst,
            "warmup": a.warmup, "ms_per_step": 1e3 * t_all / max(1, nst), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"1 env per host core x {cores} cores, {a.motion} imitation, random-action rollout "
                                   "(CPU restatement of the reference step; mujoco-py/MuJoCo 2.0 unavailable)",
                       "envs": cores, "reward_mode": a.reward_mode},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} env-steps per core per step x {nst} steps"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def redirect_stdout_to_stderr():
    """stdout must carry exactly one JSON line: send whatever any library prints to fd 1 while we run (NCCL's
    version banner, debug lines) to stderr.  Returns the saved descriptor of the real stdout (None if stderr is
    not usable and nothing was changed)."""
    saved = None
    try:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
    except OSError:
        if saved is not None:
            os.close(saved)
        saved = None
    return saved


def emit_json(line, saved_fd):
    sys.stdout.flush()
    if saved_fd is not None:
        os.write(saved_fd, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


def restore_stdout(saved_fd):
    if saved_fd is not None:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)


def run_ours(a):
    import torch
    import torch.distributed as dist
    from deepmimic_mujoco_b200.dist import RecordGather
    from deepmimic_mujoco_b200.env import DPVecEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: everything any library prints to fd 1 while we run (NCCL's version
    # banner, debug lines) is sent to stderr, and the JSON line is written to the saved descriptor at the end
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    real_stdout = redirect_stdout_to_stderr()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E = a.envs_per_gpu
    n_global = E * world
    env = DPVecEnv(E, motions=(a.motion,), device=dev, seed=0, first_env_id=rank * E, reward_mode=a.reward_mode,
                   auto_reset=True)
    sim = env.sim
    env.reset()
    gather = RecordGather(sim.rec, n_global) if world > 1 else None
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    pool = torch.rand(16, E, sim.nu, device=dev, generator=g) - 0.5          # U(-0.5, 0.5) actions
    flush = None if a.no_flush else torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    K, W = a.steps, max(a.warmup, 3)

    def one_step(i):
        env.step(pool[i % 16])
        if gather is not None:
            gather()

    for i in range(W):
        one_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- device-timed region: K steps, per-step CUDA events on the launching stream, L2 flushed between steps
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    evk = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for i in range(K):
        if flush is not None:
            flush.zero_()
        ev0[i].record()
        env.step(pool[i % 16])
        evk[i].record()
        if gather is not None:
            gather()
        ev1[i].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_dev = sum(s.elapsed_time(e) for s, e in zip(ev0, ev1)) * 1e-3
    t_kernel = sum(s.elapsed_time(e) for s, e in zip(ev0, evk)) * 1e-3
    # ---- end-to-end: host (pinned) actions in, record out, every step, through DPVecEnv.step
    h_act = [torch.empty(E, sim.nu, dtype=torch.float32).pin_memory() for _ in range(4)]
    for b in h_act:
        b.copy_(torch.rand(E, sim.nu) - 0.5)
    d_act = torch.empty(E, sim.nu, dtype=torch.float32, device=dev)
    # every rank's host reads the record of ITS OWN envs; the gathered [N,58] record stays on the device
    h_rec = torch.empty(E, sim.obs_dim + 2, dtype=torch.float32).pin_memory()

    def e2e_step(i):
        d_act.copy_(h_act[i % 4], non_blocking=True)
        env.step(d_act)
        if gather is not None:
            gather()
        h_rec.copy_(sim.rec, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller consumes obs/reward/done on the host

    for i in range(3):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([t_dev, t_kernel, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_kernel, t_e2e = [float(x) for x in tt.tolist()]
    if rank == 0:
        value = n_global * K / t_dev
        peak, peak_src = measured_peak_hbm()
        ach = ALG_BYTES_PER_ENV_STEP * E / (t_kernel / K) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(str(E))
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{E}-env {a.motion} imitation per GPU, random-action rollout, RK4+PGS(50), "
                                       f"reward_mode={a.reward_mode}, CoM termination + RSI auto-reset",
                           "envs_global": n_global, "envs_per_gpu": E, "parallelism": f"env-shard x{world}",
                           "collective": "nccl all_gather [N,58] f32 per step" if world > 1 else "none",
                           "l2": "no flush" if a.no_flush else "L2 flushed between timed steps (192 MiB memset, untimed)",
                           "launch": sim.launch_info()},
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "note": "compute/latency-bound fp32 kernel: ~1e3 FLOP/B, see DESIGN.md"},
                "e2e": {"value": n_global * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": world * E * sim.nu * 4,
                        "d2h_bytes_per_step": world * h_rec.numel() * 4},   # job totals over all ranks
                "gpu_launches": 2 * K,  # k_order (scheduler sort) + k_step (fused env step) per step
                "clocks": clocks}
        if not a.no_cpu_baseline and world == 1:   # the CPU baseline is reported by the single-GPU run only
            cpu = CpuRollout(1, a.motion, a.reward_mode)
            v, n, w = cpu.run(150000)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{n} env-steps of the same workload on 1 env (float64 oracle port), {w:.1f} s",
                                    "host_cores": host_cores()}
        emit_json(line, real_stdout)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    restore_stdout(real_stdout)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
