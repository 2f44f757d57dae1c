#!/usr/bin/env python3
"""bench.py -- env steps/s of the batched humanoid walk rollout (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]
                  [--envs-per-gpu E | --envs-global G --scaling strong] [--motions a,b,c]

--config selects a BASELINE.json configuration (default 2, the one the metric is quoted on; the driver's lines):
  2  4096-env walk imitation per GPU, weak scaling            3  16384-env spinkick, early termination, 1 GPU
  4  65536-env walk = 8192 envs per GPU x 8 (weak); with --scaling strong --envs-global 65536 the same global
     batch on 1/2/4/8 GPUs                                    5  mixed walk/dance_b/spinkick, clip = env index % 3,
                                                                 per-env RSI phase, 4096 envs per GPU (32768 on 8)

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
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--envs-global", type=int, default=None)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--motions", default=None, help="comma-separated clip names; env i uses clip i %% len")
    ap.add_argument("--motion", default=None, help="single clip (same as --motions with one name)")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--term-mode", type=int, default=0, help="1 = add the DeepMimic fall-contact termination rule")
    ap.add_argument("--sync-gather", action="store_true", help="all-gather on the compute stream (no overlap)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="p2p: all-gather fused into the step kernel over NVLink peer memory; nccl: ncclAllGather on a side stream")
    ap.add_argument("--nccl-ctas", type=int, default=0, help="NCCL max_ctas for the all-gather (0 = NCCL default)")
    ap.add_argument("--gather-depth", type=int, default=4, help="outstanding all-gathers (record buffers) in the overlapped form")
    ap.add_argument("--reward-mode", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--e2e-memcpy", action="store_true",
                    help="end-to-end number through cudaMemcpyAsync + step + cudaMemcpyAsync only (skip the host-mapped step)")
    a = ap.parse_args()
    # BASELINE.json configs[1..4]: (envs per GPU, clips, seed)
    envs, motions, seed = {2: (4096, "walk", 0), 3: (16384, "spinkick", 1), 4: (8192, "walk", 0),
                           5: (4096, "walk,dance_b,spinkick", 2)}[a.config]
    a.motions = [m for m in (a.motions or a.motion or motions).split(",") if m]
    a.motion = a.motions[0]
    a.seed = seed if a.seed is None else a.seed
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.scaling == "strong":
        g = a.envs_global or envs * (8 if a.config in (4, 5) else 1)
        if g % world:
            raise SystemExit("--envs-global must be a multiple of the number of GPUs")
        a.envs_global, a.envs_per_gpu = g, g // world
    else:
        a.envs_per_gpu = a.envs_per_gpu or (a.envs_global // world if a.envs_global else envs)
        a.envs_global = a.envs_per_gpu * world
    return a


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


def _worker_init(motions, reward_mode):
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
    motions = [motions] if isinstance(motions, str) else list(motions)
    aux = compute_ref_aux(motions) if reward_mode == 4 else None
    mcs, keep = make_mocap_struct(load_motions(motions), aux)
    L = po.lib()
    e = po.DmoEnv()
    L.dmo_env_init(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), 1234, wid, wid % len(motions))
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

    def __init__(self, cores, motions, reward_mode):
        import multiprocessing as mp
        self.cores = cores
        self.pool = None
        if cores == 1:
            _worker_init(motions, reward_mode)
        else:
            self.pool = mp.get_context("fork").Pool(cores, initializer=_worker_init, initargs=(motions, reward_mode))

    def run(self, nsteps):
        res = [_worker_run(nsteps)] if self.pool is None else self.pool.map(_worker_run, [nsteps] * self.cores, chunksize=1)
        total = sum(r[0] for r in res)
        wall = max(r[1] for r in res)
        return total / wall, total, wall

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def python_driven_cpu(motion, reward_mode, nsteps):
    """the same float64 port driven from Python, one ctypes call per env step (what a
    mujoco-py user pays per step), random actions from numpy, reset on done.  Returns env-steps/s."""
    import ctypes as C
    import numpy as np
    _worker_init(motion, reward_mode)
    L, m, cfg, mcs, e = _W["L"], _W["m"], _W["cfg"], _W["mcs"], _W["e"]
    import oracle.pyoracle as po
    rng = np.random.default_rng(0)
    acts = rng.uniform(-0.5, 0.5, size=(256, m.nu))
    obs = np.zeros(256); rew = C.c_double()
    t0 = time.perf_counter()
    for t in range(nsteps):
        L.dmo_env_step(C.byref(m), C.byref(cfg), C.byref(mcs), C.byref(e), po.dptr(acts[t & 255]), po.dptr(obs), C.byref(rew))
    return nsteps / (time.perf_counter() - t0)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = 1000  # env-steps per core per bench "step": a bounded sample of the same workload
    cpu = CpuRollout(cores, a.motions, a.reward_mode)
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
            "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE config {a.config}: 1 env per host core x {cores} cores, {'+'.join(a.motions)} imitation, random-action rollout "
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
    import numpy as np
    import torch
    import torch.distributed as dist
    from deepmimic_mujoco_b200.dist import PeerRecordGather, RecordGather, mixed_clip_ids
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
        # The step kernel is one persistent CTA per SM that owns the SM's whole shared memory, so NCCL's CTAs cannot share an
        # SM with it: when the gather of step t is still waiting for the slowest rank, its CTAs sit on 16-32 SMs in front
        # of the CTAs of step t + 1, which couples the ranks (8 GPUs: the kernel takes 0.468 ms on EVERY rank instead of
        # 0.445).  Two knobs were measured against that (profiles/r2_c22_nccl_cta_experiment.txt) and are OFF by default:
        # --nccl-ctas k (NCCL max_ctas; 1 CTA makes the gather itself take 1 ms, 4 CTAs 0.6 ms) and DMB_RESERVE_SMS=k
        # (dmb_create leaves k SMs without a persistent CTA; costs 1.5 % at 8192 envs per GPU).
        opts = None
        if a.nccl_ctas > 0:
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.min_ctas = 1
                opts.config.max_ctas = a.nccl_ctas
            except Exception:
                opts = None
        if opts is not None:
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        else:
            dist.init_process_group("nccl", device_id=dev)
    E = a.envs_per_gpu
    n_global = E * world
    first = rank * E
    nclip = len(a.motions)
    clip_ids = mixed_clip_ids(first, first + E, nclip) if nclip > 1 else None   # clip = global env index % nclip
    env = DPVecEnv(E, motions=tuple(a.motions), device=dev, seed=a.seed, first_env_id=first, reward_mode=a.reward_mode,
                   auto_reset=True, clip_ids=clip_ids, term_mode=a.term_mode, rec_depth=max(2, a.gather_depth))
    sim = env.sim
    env.reset()
    D = max(2, a.gather_depth)
    p2p = world > 1 and a.gather == "p2p" and not a.sync_gather
    gather = RecordGather(sim.rec, n_global, depth=D) if world > 1 and not p2p else None
    # fused form: the step kernel itself stores every record row into all ranks' gathered buffers (NVLink peer memory)
    # and signals them; the consumer's wait for step t - 2 is folded into the kernel of step t
    peer, p2p_error = None, None
    if p2p:
        try:
            peer = PeerRecordGather(sim, n_global, first, depth=max(D, 6))
        except Exception as ex:          # no CUDA IPC / peer access on this box: every rank fails alike -> NCCL form
            p2p_error = f"{type(ex).__name__}: {ex}"
            p2p = False
            gather = RecordGather(sim.rec, n_global, depth=D)
    LAG = 1   # the wait for step t - 2 is folded into the kernel of step t; 6 buffers >= 2 * 2 + 1 (see PeerRecordGather)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    pool = torch.rand(16, E, sim.nu, device=dev, generator=g) - 0.5          # U(-0.5, 0.5) actions
    flush = None if a.no_flush else torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    K, W = a.steps, max(a.warmup, 3)

    # One step = the fused env-step kernel + (N > 1) the NCCL all-gather of its [E, 58] record.  The gather of step
    # t runs on a side stream and overlaps the following steps (D record buffers): the compute stream joins it right
    # after the kernel of step t + D - 1, so every gather lies inside some step's timed interval (the last D - 1
    # are drained and timed after the loop).
    def one_step(i):
        if peer is not None:
            peer.arm()
            if i >= LAG + 1:
                peer.wait(in_next_step=True)   # step i - LAG - 1 ... folded into this step's kernel
        env.step(pool[i % 16])
        if gather is not None:
            if a.sync_gather:
                gather(sim.rec)
            else:
                if i >= D - 1:
                    gather.wait()
                gather.launch(sim.rec)

    for i in range(W):
        one_step(i)
    if gather is not None:
        gather.drain()
    if peer is not None:
        peer.drain()
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
    evd = torch.cuda.Event(enable_timing=True)
    launches0 = sim.kernel_launches()
    for i in range(K):
        if flush is not None:
            flush.zero_()
        ev0[i].record()
        if peer is not None:
            peer.arm()
            if peer._t - peer._waited > LAG + 1:
                # every rank's rows of step i - LAG - 1 must have arrived before this step's kernel completes: one
                # thread of the kernel polls the flag at its start (the wait costs no launch of its own)
                peer.wait(in_next_step=True)
        env.step(pool[i % 16])
        evk[i].record()
        if gather is not None:
            if a.sync_gather:
                gather(sim.rec)
            else:
                if i >= D - 1:
                    gather.wait()          # gather of step i-D+1, launched D-1 steps ago
                gather.launch(sim.rec)
        ev1[i].record()
    if gather is not None:
        gather.drain()                     # the last gather has nothing to hide behind: it is timed on its own
    if peer is not None:
        peer.drain()
    evd.record()
    launches = sim.kernel_launches() - launches0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_dev = (sum(s.elapsed_time(e) for s, e in zip(ev0, ev1)) + ev1[-1].elapsed_time(evd)) * 1e-3
    t_kernel = sum(s.elapsed_time(e) for s, e in zip(ev0, evk)) * 1e-3
    # ---- end-to-end: host (pinned) actions in, record out, every step, through DPVecEnv.step
    h_act = [torch.empty(E, sim.nu, dtype=torch.float32).pin_memory() for _ in range(4)]
    for b in h_act:
        b.copy_(torch.rand(E, sim.nu) - 0.5)
    d_act = torch.empty(E, sim.nu, dtype=torch.float32, device=dev)
    # every rank's host reads the record of ITS OWN envs; the gathered [N,58] record stays on the device
    h_rec = torch.empty(E, sim.obs_dim + 2, dtype=torch.float32).pin_memory()

    def all_ranks(ok):                         # every rank takes the same branch (time_e2e holds a barrier)
        if world == 1:
            return bool(ok)
        f = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        return bool(f.item())

    # Host-mapped buffers (DPVecEnv.step(..., rec_host=) / step_host): the step kernel itself loads the action rows from,
    # and / or stores the record rows to, pinned device-mapped host memory over PCIe instead of a cudaMemcpyAsync on
    # that side.  (With the NCCL fallback gather the record has to stay on the device: memcpy form only.)
    m_act = m_rec = None
    e2e_note = None
    if not a.e2e_memcpy and gather is None:
        try:
            m_act = [sim.alloc_host((E, sim.nu)) for _ in range(4)]
            for b, src in zip(m_act, h_act):
                b.array[:] = src.numpy()
            m_rec = sim.alloc_host((E, sim.obs_dim + 2))
        except Exception as ex:                # the memcpy form stands
            e2e_note = f"host-mapped buffers unavailable ({type(ex).__name__}: {ex})"
            m_rec = None
    mapped_ok = all_ranks(m_rec is not None)
    if not mapped_ok and e2e_note is None and not a.e2e_memcpy and gather is None:
        e2e_note = "host-mapped buffers unavailable on another rank"

    def make_e2e_step(act_mapped, rec_mapped):
        def f(i):
            if act_mapped:
                act = m_act[i % 4]
            else:
                d_act.copy_(h_act[i % 4], non_blocking=True)
                act = d_act
            if peer is not None:
                peer.arm()
            env.step(act, rec_host=m_rec if rec_mapped else None)
            if peer is not None:
                peer.wait()                    # the caller consumes this step's gathered record: no overlap here
            if gather is not None:
                gather.launch(sim.rec)
                gather.wait()                  # the caller consumes this step's gathered record: no overlap here
            if not rec_mapped:
                h_rec.copy_(sim.rec, non_blocking=True)
            torch.cuda.current_stream().synchronize()   # the caller consumes obs/reward/done on the host
        return f

    def time_e2e(step_fn):
        for i in range(3):
            step_fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(K):
            step_fn(i)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        if world > 1:                          # one decision for the job: the slowest rank's time
            tm = torch.tensor([t], device=dev, dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            t = float(tm[0])
        return t

    E2E_FORMS = [("memcpy", False, False)]
    if mapped_ok:
        E2E_FORMS += [("memcpy_action+host_mapped_record", False, True), ("host_mapped_action+memcpy_record", True, False),
                      ("host_mapped", True, True)]
    e2e_times = {}
    for name, am, rm in E2E_FORMS:
        try:
            t = time_e2e(make_e2e_step(am, rm))
            ok = True
            if rm:                             # the host record of the last step must be what the device buffers hold
                od = sim.obs_dim
                ok = (np.array_equal(m_rec.array[:, :od], sim.obs.cpu().numpy())
                      and np.array_equal(m_rec.array[:, od], sim.reward.cpu().numpy())
                      and np.array_equal(m_rec.array[:, od + 1] != 0, sim.done.cpu().numpy() != 0))
        except Exception as ex:
            if world > 1 or name == "memcpy":  # other ranks are inside a collective / nothing to fall back to
                raise
            e2e_note = f"{name} form failed ({type(ex).__name__}: {ex})"
            continue
        if all_ranks(ok):
            e2e_times[name] = t
        else:
            e2e_note = f"{name}: host record differs from the device buffers on some rank (form dropped)"
    e2e_mode = min(e2e_times, key=e2e_times.get)    # the fastest validated form is the end-to-end number
    t_e2e = e2e_times[e2e_mode]
    clocks = sampler.stop() if sampler else None
    # ---- N = 1 gym surface (the reference's own use: trpo.py with one env): DPEnv.step latency, numpy in / out
    gym_sps = None
    if rank == 0 and world == 1 and a.config == 2:
        try:
            from deepmimic_mujoco_b200.env import DPEnv
            e1 = DPEnv(motion=a.motion, device=dev, seed=0)
            e1.reset()
            acts = [e1.action_space.sample() for _ in range(64)]
            for i in range(30):
                if e1.step(acts[i % 64])[2]:
                    e1.reset()
            torch.cuda.synchronize()
            tg = time.perf_counter()
            ng = 300
            for i in range(ng):
                if e1.step(acts[i % 64])[2]:
                    e1.reset(); e1.reset_model_init()
            gym_sps = ng / (time.perf_counter() - tg)
            e1.close()
        except Exception as ex:   # diagnostics only
            gym_sps = f"failed: {ex}"
    spread = None
    if world > 1:
        tmin = torch.tensor([t_dev, t_kernel], device=dev, dtype=torch.float64)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        tt = torch.tensor([t_dev, t_kernel], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_kernel = [float(x) for x in tt.tolist()]
        # how much of the multi-GPU step time is rank-to-rank variation of the kernel itself (different envs on
        # every rank: the slowest env of the slowest rank sets the pace) and how much is the collective
        spread = {"kernel_ms_per_step_min_rank": 1e3 * float(tmin[1]) / K, "kernel_ms_per_step_max_rank": 1e3 * t_kernel / K,
                  "step_ms_min_rank": 1e3 * float(tmin[0]) / K, "step_ms_max_rank": 1e3 * t_dev / K}
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
        wl = "+".join(a.motions)
        rec_w = sim.obs_dim + 2
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"BASELINE config {a.config}: {E}-env {wl} imitation per GPU, random-action rollout, "
                                       f"RK4+PGS(50), reward_mode={a.reward_mode}, CoM termination"
                                       f"{' + fall-contact rule' if a.term_mode == 1 else ''} + RSI auto-reset"
                                       f"{', clip = env index % ' + str(nclip) if nclip > 1 else ''}",
                           "baseline_config": a.config, "envs_global": n_global, "envs_per_gpu": E,
                           "parallelism": f"env-shard x{world}",
                           "collective": (f"all-gather of the [N,{rec_w}] f32 record fused into the step kernel: peer stores over NVLink "
                                          f"+ arrival flags, consumer waits for step t-{LAG + 1} inside the step kernel ({max(D, 6)} buffers); no NCCL kernel") if p2p else
                                         (f"nccl all_gather [N,{rec_w}] f32 per step, "
                                          + ("on the compute stream" if a.sync_gather else
                                             f"side stream, overlapped with the following steps ({D} record buffers, "
                                             f"{a.nccl_ctas or 'default'} NCCL CTA(s), {os.environ.get('DMB_RESERVE_SMS', '0')} SM(s) left free)"))
                                         if world > 1 else "none",
                           "l2": "no flush" if a.no_flush else "L2 flushed between timed steps (192 MiB memset, untimed)",
                           "launch": sim.launch_info()},
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": traffic, "peak_source": peak_src,
                             "note": "compute/latency-bound fp32 kernel: ~1e3 FLOP/B, see DESIGN.md"},
                # e2e: host (pinned) actions in, record out, every step, through DPVecEnv.step.  Each side is either a
                # cudaMemcpyAsync around the step or host-mapped (the step kernel reads / writes the pinned host rows
                # itself over PCIe); every form is timed (forms: env-steps/s each) and the fastest one whose host record
                # matched the device buffers is reported
                "e2e": {"value": n_global * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": world * E * sim.nu * 4,
                        "d2h_bytes_per_step": world * h_rec.numel() * 4,   # job totals over all ranks
                        "mode": e2e_mode, "forms": {k: n_global * K / v for k, v in e2e_times.items()}},
                # our kernels in the timed region, counted by the library: k_step (fused env step) every step + k_order
                # (scheduler sort; every step with dynamic pulling, every 8th step when one round holds every env)
                "gpu_launches": launches,
                "clocks": clocks}
        if spread is not None:
            line["rank_spread"] = spread
        if e2e_note is not None:
            line["e2e"]["note"] = e2e_note
        if p2p_error is not None:
            line["config"]["p2p_gather_unavailable"] = p2p_error
        if world > 1:
            line["nvlink"] = {"gather_bytes_in_per_rank_per_step": (world - 1) * E * rec_w * 4,
                              "gather_bytes_out_per_rank_per_step": E * rec_w * 4}
        if gym_sps is not None:
            line["gym_surface_n1"] = {"value": gym_sps, "unit": "env-steps/s",
                                      "what": "DPEnv.step (N = 1, numpy in/out; one launch per step, action and record in pinned host memory)"}
        if not a.no_cpu_baseline and world == 1:   # the CPU baseline is reported by the single-GPU run only
            cpu = CpuRollout(1, a.motions, a.reward_mode)
            v, n, w = cpu.run(150000)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{n} env-steps of the same workload on 1 env (float64 oracle port, C loop), {w:.1f} s",
                                    "host_cores": host_cores()}
            pv = python_driven_cpu(a.motions[0], a.reward_mode, 20000)
            line["cpu_baseline"]["python_driven"] = {"value": pv, "unit": UNIT,
                                                     "what": "same port, one ctypes call per env step (mujoco-py-like call overhead)"}
            cpu.close()
        emit_json(line, real_stdout)
    if peer is not None:
        peer.close()
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
