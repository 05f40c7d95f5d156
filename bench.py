#!/usr/bin/env python
"""bench.py -- Allsteps-v0 batched MDP step on B200 (and the reference's CPU implementation beside it).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path, N ranks via torchrun
    python bench.py --impl reference --steps K --warmup W      # the reference algorithm on the host cores

A "step" is one pass of the hot path (direct_rl_env.py:351-375 of the reference: episode counter, pass 1, dones,
rewards, masked reset incl. the PhysX start-pose rows, pass 2, observations) over one batch of synthetic
articulation state.  Workload at every N: 1,048,576 envs per GPU (weak scaling; envs are
sharded by env id, no collective on the step path; step statistics are all-reduced over NCCL off the step path).

The timed state stays on SURVEY 8(d)'s distribution: the input sets form a cycle, set k generated from the MDP state k
steps of the cycle lead to, and the MDP state is rewound every `--input-sets` steps INSIDE the timed region
(allsteps_isaaclab_b200/workload.py); the advance / reset rates that result are printed in the line.

JSON keys: `value` = env-steps/s with inputs resident in HBM (CUDA events, max over ranks); `e2e` = the same metric
through the public API with HOST buffers (pinned host -> device copy of every step's inputs and device -> host
read of its results inside the timed region); `roofline` = SURVEY 8(d)'s 652 algorithmic bytes per env-step over the
WHOLE step's time against the measured HBM copy peak (`roofline.kernel` = the dominant kernel alone, timed live by
CUDA events on its stream); `cpu_baseline` = the CPU oracle port on a bounded sample.  First-class blocks for the
other BASELINE.json configs: `c2_4096` (rl_games' default scale), `c3_65536_grid`, `c5_reset_heavy`, and at N > 1
`c4` (1,048,576 envs SPLIT over the N ranks, promotion on the global mean over NVLink peer memory, with a
peer == NCCL == single-handle bit-identity self-check); `three_call` = the pass1 / reset / pass2 path the DirectRLEnv
hooks use under PhysX; `isaac_layout` = the fused step on the (N,13) / (N,B,13) views Isaac Lab hands out.
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
sys.path.insert(0, ROOT)

B_ALG = 652  # algorithmic bytes per env-step (BASELINE.md section 4 / SURVEY.md section 8d)
ENVS_PER_GPU = 1 << 20
CPU_SAMPLE_ENVS = 1 << 16
C4_GLOBAL_ENVS = 1 << 20
METRIC = "MDP env-steps/sec"
UNIT = "env-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--input-sets", type=int, default=16,
                    help="length of the cycle of synthetic states (each generated from the MDP state it meets; the MDP "
                         "state is rewound when the cycle restarts)")
    ap.add_argument("--no-extra-blocks", action="store_true",
                    help="skip c2/c3/c5/three_call/isaac_layout (N=1) and c4 (N>1): headline numbers only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 24)")
    ap.add_argument("--stats-interval", type=int, default=16, help="all-reduce step statistics every k steps")
    ap.add_argument("--global-promotion", nargs="?", const="peer", default=None, choices=["peer", "nccl"],
                    help="N > 1: apply the promotion rule ENV:471 to the mean over ALL shards every step (config 4). "
                         "peer (default when given): one exchange kernel over NVLink peer memory; nccl: as_fold_stats + "
                         "NCCL all-reduce + as_finish_step(global). Without the flag promotion is shard-local.")
    # other BASELINE.json configs (the default flags are the headline workload)
    ap.add_argument("--fall-fraction", type=float, default=0.02, help="fraction of envs dying per step (0.3 = config 5)")
    ap.add_argument("--intended-regen", action="store_true", help="regenerate stones of reset envs past S/2 (extension)")
    ap.add_argument("--grid-bins", type=int, default=0, help="pitch x yaw grid curriculum with B x B bins (config 3)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram bytes per launch of the fused step kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return d
    return None


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """`nvidia-smi -lms 100` running in the background during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)  # let the first sample land before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            self.lines = [ln for ln in out.splitlines() if ln.strip()]

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            s = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        busy = sorted(x for x in sm if x > 0.5 * (max(mx) if mx else 0)) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def run_cpu_port(num_envs: int, steps: int, warmup: int, seed: int = 1234, period: int = 4):
    """Times the CPU oracle port (the reference's algorithm, op for op in torch) on all host cores.  Same cycle of
    synthetic states as the CUDA arm: set k is generated from the MDP state k steps lead to, the oracle's MDP buffers
    are rewound (outside the per-step timers) when the cycle restarts.  The port leaves out the six debug clones of
    ENV:257-266, i.e. it is slightly FASTER than the reference itself."""
    import copy

    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from scenario import Scenario, install_mdp_state
    from oracle import allsteps_oracle as ao

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc = Scenario(num_envs, seed=seed)
    # (level 0: the stone draws have zero-width ranges, ENV:126-133 -- no table of uniforms needed)
    orc = ao.AllstepsOracle(sc.cfg, num_envs, sc.env_origins, sc.joint_limits, sc.body_indices, None)
    install_mdp_state(orc, sc.initial_mdp_state())
    mirror_u = torch.rand(num_envs, generator=sc.gen)
    noise_u = torch.rand(num_envs, 21, generator=sc.gen)
    state_keys = ("curr_target_index", "prev_target_index", "next_target_index", "swing_leg", "target_reach_count",
                  "episode_length_buf", "curriculum", "potentials", "old_potentials")
    snap = {k: getattr(orc, k).clone() for k in state_keys}
    pool = []
    for _ in range(period):
        phys = sc.physics(orc.steps_pos, orc.curr_target_index, orc.swing_leg)
        pool.append(phys)
        orc.step(phys, phys["actions"], mirror_u, noise_u, None)
    times, adv = [], 0
    for i in range(warmup + steps):
        if i % period == 0:
            for k in state_keys:
                setattr(orc, k, snap[k].clone())
        phys = pool[i % period]
        t0 = time.perf_counter()
        orc.step(phys, phys["actions"], mirror_u, noise_u, None)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    return {"value": num_envs * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores,
            "threads": torch.get_num_threads()}


def run_eager_cuda_port(torch, dev, num_envs: int, steps: int, warmup: int, seed: int = 1234, period: int = 4):
    """"What users get today": the reference's algorithm as eager torch ops ON THE GPU (the oracle port with CUDA
    tensors -- about 340 small kernels and the reference's own device->host syncs per pass), same cycle of synthetic
    states.  Wall time per step (the path synchronises by itself: `.nonzero()`, the camera follow of ENV:323-324)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from scenario import install_mdp_state
    from allsteps_isaaclab_b200 import synthetic as syn
    from allsteps_isaaclab_b200.config import AllstepsCfg
    from oracle import allsteps_oracle as ao

    cfg = AllstepsCfg()
    with torch.device(dev):
        origins = syn.env_origins_grid(num_envs, cfg.env_spacing).to(dev)
        limits = syn.joint_limits_tensor(cfg).to(dev)
        orc = ao.AllstepsOracle(cfg, num_envs, origins, limits, (0, 1, 2), None)
        st0 = {k: v.to(dev) for k, v in syn.random_mdp_state(cfg, num_envs, torch.Generator().manual_seed(seed)).items()}
        install_mdp_state(orc, st0)
        gen = torch.Generator(device=dev).manual_seed(seed)
        mirror_u = torch.rand(num_envs, generator=gen, device=dev)
        noise_u = torch.rand(num_envs, 21, generator=gen, device=dev)
        keys = ("curr_target_index", "prev_target_index", "next_target_index", "swing_leg", "target_reach_count",
                "episode_length_buf", "curriculum", "potentials", "old_potentials")
        snap = {k: getattr(orc, k).clone() for k in keys}
        pool = []
        for _ in range(period):
            phys = syn.random_physics_state(cfg, orc.steps_pos, orc.curr_target_index, orc.swing_leg, gen)
            pool.append(phys)
            orc.step(phys, phys["actions"], mirror_u, noise_u, None)
        times = []
        for i in range(warmup + steps):
            if i % period == 0:
                for k in keys:
                    setattr(orc, k, snap[k].clone())
            phys = pool[i % period]
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            orc.step(phys, phys["actions"], mirror_u, noise_u, None)
            torch.cuda.synchronize(dev)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": num_envs * len(times) / total, "us_per_step": 1e6 * total / len(times), "envs": num_envs}


def main_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (it is pure Python/torch and cannot be
    installed on the GPU box, so the bit-identical oracle port stands in -- kind "port"), all host threads, on the
    same config as the CUDA arm: 1,048,576 envs per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    envs = args.envs_per_gpu
    # ~0.5 us of host time per env-step on 16 cores: K steps of 1M envs stay within a few minutes up to K ~ 200
    while envs > 4096 and envs * (args.steps + max(args.warmup, 3)) * 1.0e-6 > 240.0:
        envs //= 2
    r = run_cpu_port(envs, args.steps, max(args.warmup, 3))
    frac = "" if envs == args.envs_per_gpu else f" (1/{args.envs_per_gpu // envs} of the workload: bounded sample)"
    sample = (f"{envs} envs per step{frac}, {args.steps} steps; the port omits the reference's six debug clones "
              "(ENV:257-266), so it is slightly faster than the reference")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Allsteps-v0 fused MDP step, {args.envs_per_gpu} envs per GPU (reference arm: the "
                               "reference's algorithm on the host cores of rank 0)", "envs_per_gpu": args.envs_per_gpu,
                   "envs_per_step": envs},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- CUDA arm
def time_cycle(torch, wl, steps, warmup, dist=None, after_step=None):
    """K steps of the workload cycle between CUDA events (the rewind at each cycle start is inside)."""
    mdp = wl.mdp
    dev = mdp.device
    wl.rewind()
    for _ in range(warmup):
        wl.step()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize(dev)
    launches0 = mdp.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        wl.step()
        if after_step is not None:
            after_step(i)
    e1.record()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1), mdp.launch_count - launches0


def time_cycle_graph(torch, wl, steps, warmup):
    """Same loop with every set's step captured once as a CUDA graph and replayed (the rewind stays an API call)."""
    mdp, dev = wl.mdp, wl.mdp.device
    wl.rewind()
    graphs = [mdp.capture_step(v, d["actions"], wl.out) for v, d in wl.sets]
    per_replay = graphs[0].kernels_per_replay

    def run(n):
        for _ in range(n):
            k = wl.j % wl.period
            if k == 0 and wl.j > 0:
                wl.rewind()
            graphs[k].replay()
            wl.j += 1

    wl.rewind()
    run(warmup)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1), steps * per_replay


def time_kernel_only(torch, wl, steps):
    """Average duration of the dominant kernel alone, k_step<fused>: the library records a pair of CUDA events on the
    launching stream immediately around that launch (as_set_timing_events)."""
    from allsteps_isaaclab_b200 import _cabi

    mdp, dev = wl.mdp, wl.mdp.device
    pairs = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b.record()  # creates the underlying cudaEvent_t
        pairs.append((a, b))
    torch.cuda.synchronize(dev)
    wl.rewind()
    for i in range(steps):
        a, b = pairs[i]
        _cabi.check(mdp.lib.as_set_timing_events(mdp.handle, a.cuda_event, b.cuda_event), "as_set_timing_events")
        wl.step()
    _cabi.check(mdp.lib.as_set_timing_events(mdp.handle, None, None), "as_set_timing_events")
    torch.cuda.synchronize(dev)
    ms = sorted(a.elapsed_time(b) for a, b in pairs)
    return sum(ms) / len(ms), ms[len(ms) // 2]


def time_per_step(torch, wl, steps):
    """Every step between its own pair of CUDA events (median and p5 / p95 of the step time)."""
    dev = wl.mdp.device
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    wl.rewind()
    for _ in range(3):
        wl.step()
    torch.cuda.synchronize(dev)
    ev[0].record()
    for i in range(steps):
        wl.step()
        ev[i + 1].record()
    torch.cuda.synchronize(dev)
    us = sorted(1e3 * ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    pick = lambda q: us[min(len(us) - 1, int(q * len(us)))]  # noqa: E731
    return {"p5": pick(0.05), "p50": pick(0.5), "p95": pick(0.95), "steps": steps,
            "note": "one CUDA-event pair per step; the steps that start a cycle include the state rewind"}


def time_mirror(torch, mdp, cfg, rows, iters, peak):
    """the mirror-symmetry augmentation of `A2CAgentSymmetry.play_steps`
    (learning/a2c_ppo_mirroring.py:32-34 -> ENV:611-660) on `rows` = horizon x envs rows: vstack((x, mirrored(x))) for
    obses (rows,59), actions and mus (rows,21) -- one launch of k_mirror_batch against the reference's own torch ops
    (clone + three fancy-index assignments + vstack per tensor) run eagerly on the same GPU."""
    from allsteps_isaaclab_b200 import symmetry
    from oracle import allsteps_oracle as ao  # (baseline leg only)

    dev = mdp.device
    g = torch.Generator(device=dev).manual_seed(5)
    obs = torch.randn(rows, 59, device=dev, generator=g)
    act = torch.randn(rows, 21, device=dev, generator=g)
    mus = torch.randn(rows, 21, device=dev, generator=g)
    tabs = (cfg.right_joint_indices, cfg.left_joint_indices, cfg.negation_joint_indices)

    def ours():
        return symmetry.mirror_batch(mdp, obs, act, mus)

    def eager():
        return (ao.symmetric_states(obs, *tabs, "obs"), ao.symmetric_states(act, *tabs, "actions"),
                ao.symmetric_states(mus, *tabs, "actions"))

    same = all(torch.equal(a, b) for a, b in zip(ours(), eager()))
    res = {}
    for name, f, n in (("k_mirror_batch", ours, iters), ("eager_torch", eager, max(3, iters // 4))):
        for _ in range(3):
            f()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            f()
        e1.record()
        torch.cuda.synchronize(dev)
        res[name] = e0.elapsed_time(e1) / n * 1e3
    b_alg = rows * (59 + 21 + 21) * 4 * 3  # every input element read once, written twice
    return {"rows": rows, "us": res["k_mirror_batch"], "eager_torch_us": res["eager_torch"],
            "speedup_vs_eager_torch": res["eager_torch"] / res["k_mirror_batch"],
            "algorithmic_GBps": b_alg / (res["k_mirror_batch"] * 1e-6) / 1e9,
            "roofline_frac": b_alg / (res["k_mirror_batch"] * 1e-6) / 1e9 / peak, "bit_identical_to_eager_torch": same}


def time_action(torch, dev, n, iters, peak):
    """ENV:257-274 (clamp, gain[level] * gear * action), called `decimation` = 4 times per env
    step: 84 bytes in + 84 bytes out per env; six rotating buffer pairs (1 GB at 1 M envs: nothing survives in L2)."""
    from allsteps_isaaclab_b200.mdp import AllstepsMDP

    m = AllstepsMDP(n, device=dev, seed=3)
    sets = [(-1.5 + 3.0 * torch.rand(n, 21, device=dev), torch.empty(n, 21, device=dev)) for _ in range(6)]
    for a, e in sets:
        m.apply_action(a, e)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        a, e = sets[i % len(sets)]
        m.apply_action(a, e)
    e1.record()
    torch.cuda.synchronize(dev)
    us = e0.elapsed_time(e1) / iters * 1e3
    return {"envs": n, "us_per_call": us, "algorithmic_GBps": n * 168 / (us * 1e-6) / 1e9,
            "roofline_frac": n * 168 / (us * 1e-6) / 1e9 / peak}


def time_three_call(torch, wl, origins, steps, warmup, device_reset_list=False):
    """DRL:351-375 as the DirectRLEnv hooks run it under PhysX: as_step_pass1, the host's `.nonzero()` on reset_buf
    (DRL:359, a device->host sync), as_reset on those ids, as_step_pass2.  (No PhysX here: pass 2 sees unchanged
    physics.)  Device time between events, host gaps included."""
    mdp, dev, out = wl.mdp, wl.mdp.device, wl.out
    N = mdp.num_envs
    ep_len = torch.zeros(N, dtype=torch.int64, device=dev)

    def one(v, d):
        ep_len.add_(1)
        mdp.pass1(v, d["actions"], out, episode_length=ep_len)
        if device_reset_list:  # the envs pass 1 flagged, from the list it compacted on the device: no host round trip
            mdp.reset(origins, None, out, episode_length=ep_len)
            mdp.pass2(v, out)
            return
        ids = out.dones.nonzero(as_tuple=False).squeeze(-1)  # DRL:359
        if len(ids) > 0:
            mdp.reset(origins, ids, out, episode_length=ep_len)
            mdp.pass2(v, out)
        else:
            mdp.no_reset()

    def run(n):
        for _ in range(n):
            k = wl.j % wl.period
            if k == 0 and wl.j > 0:
                wl.rewind()
            one(*wl.sets[k])
            wl.j += 1

    wl.rewind()
    run(warmup)
    torch.cuda.synchronize(dev)
    l0 = mdp.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) * 1e3
    return max(e0.elapsed_time(e1), wall), mdp.launch_count - l0


def time_e2e(torch, wl, origins, steps, warmup, zero_copy_contact=False):
    """Same metric through the public API with HOST buffers: every step copies that step's inputs from pinned host
    memory, runs the fused step, and reads the results back to pinned host memory.  Copies of step t+1 / t-1
    overlap the kernel of step t on separate streams (double-buffered device inputs and outputs).  The host holds the
    first two states of the cycle; the MDP state is rewound every two steps.

    zero_copy_contact: the two (N,1,20,3) contact matrices (59 % of the input bytes, of which the step needs 24 B per
    env) are NOT copied; the C ABI is handed the pinned host tensors themselves (device-accessible under unified
    addressing) and the contact-gather kernel fetches just the current stone's vectors across PCIe."""
    from allsteps_isaaclab_b200.mdp import PhysicsViews, StepBuffers

    mdp = wl.mdp
    dev = mdp.device
    N = mdp.num_envs
    keys = ["root_pos_w", "root_quat_w", "root_lin_vel_w", "body_pos_w", "joint_pos", "joint_vel",
            "force_matrix_right", "force_matrix_left", "actions"]
    host_sets = []
    for _, d in wl.sets[:2]:
        host_sets.append({k: d[k].cpu().pin_memory() for k in keys})
    copied = [k for k in keys if not (zero_copy_contact and k.startswith("force_matrix"))]
    h2d_bytes = sum(host_sets[0][k].numel() * host_sets[0][k].element_size() for k in copied)
    if zero_copy_contact:
        h2d_bytes += N * 2 * 32  # what the gather kernel pulls over PCIe: one 32-byte sector per foot and env, at least
    dev_in = [{k: torch.empty_like(host_sets[0][k], device=dev) for k in copied} for _ in range(2)]
    if zero_copy_contact:
        views = [[PhysicsViews.from_dict({**dev_in[b], "force_matrix_right": hs["force_matrix_right"],
                                          "force_matrix_left": hs["force_matrix_left"]}, origins)
                  for hs in host_sets] for b in range(2)]
    else:
        views = [[PhysicsViews.from_dict(dev_in[b], origins)] * len(host_sets) for b in range(2)]
    outs = [wl.out, StepBuffers(N, dev)]
    host_out = [{"obs": torch.empty(N, 59).pin_memory(), "reward": torch.empty(N).pin_memory(),
                 "terminated": torch.empty(N, dtype=torch.bool).pin_memory(),
                 "time_out": torch.empty(N, dtype=torch.bool).pin_memory()} for _ in range(2)]
    d2h_bytes = sum(t.numel() * t.element_size() for t in host_out[0].values())
    s_in, s_out, s_main = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    out_ready = [torch.cuda.Event() for _ in range(2)]
    out_free = [torch.cuda.Event() for _ in range(2)]

    def h2d(i):
        b = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(in_free[b])
            for k in copied:
                dev_in[b][k].copy_(host_sets[i % len(host_sets)][k], non_blocking=True)
            in_ready[b].record(s_in)

    def run(i):
        b = i % 2
        s_main.wait_event(in_ready[b])
        s_main.wait_event(out_free[b])
        if i % 2 == 0:
            wl.rewind()  # the two host-resident states belong to the first two steps of the cycle
        mdp.step(views[b][i % len(host_sets)], dev_in[b]["actions"], outs[b])
        in_free[b].record(s_main)
        out_ready[b].record(s_main)

    def d2h(i):
        b = i % 2
        with torch.cuda.stream(s_out):
            s_out.wait_event(out_ready[b])
            host_out[b]["obs"].copy_(outs[b].obs, non_blocking=True)
            host_out[b]["reward"].copy_(outs[b].reward, non_blocking=True)
            host_out[b]["terminated"].copy_(outs[b].terminated, non_blocking=True)
            host_out[b]["time_out"].copy_(outs[b].time_out, non_blocking=True)
            out_free[b].record(s_out)

    for b in range(2):
        in_free[b].record(s_main)
        out_free[b].record(s_main)
    warmup += warmup % 2  # keep the timed region on the cycle's phase
    total = warmup + steps
    t_start = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d(0)
    for i in range(total):
        if i == warmup:
            torch.cuda.synchronize(dev)
            h2d(i)  # re-issue: the timed region starts with the copy of its first step's inputs
            e0.record(s_main)
            t_start = time.perf_counter()
        if i + 1 < total:
            h2d(i + 1)
        run(i)
        d2h(i)
    s_main.wait_stream(s_out)
    e1.record(s_main)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t_start
    ms = max(e0.elapsed_time(e1), wall * 1e3)
    checksum = float(host_out[(total - 1) % 2]["reward"].sum())
    return ms, h2d_bytes, d2h_bytes, checksum


def selfcheck_global_promotion(torch, dist, dev, rank, world, per=8192, steps=10, seed=17):
    """BASELINE config 4 correctness inside the bench run (the driver's test box has one GPU): every rank steps its
    env-id shard three ways from the same seeded state -- global promotion over NVLink peer memory, the same through
    as_fold_stats + NCCL all-reduce + as_finish_step(global), and rank 0 additionally ONE handle with all envs --
    and all outputs and MDP state must be bit-identical.  The state sits near the promotion threshold so that
    promotions actually happen."""
    from allsteps_isaaclab_b200 import synthetic as syn
    from allsteps_isaaclab_b200.config import AllstepsCfg
    from allsteps_isaaclab_b200.mdp import AllstepsMDP, PhysicsViews, StepBuffers

    cfg = AllstepsCfg()
    N = per * world
    sl = slice(rank * per, (rank + 1) * per)
    g0 = torch.Generator().manual_seed(seed)            # same on every rank
    st0 = syn.random_mdp_state(cfg, N, g0)
    st0["curr_target_index"] = torch.randint(11, 20, (N,), generator=g0)
    origins_all = syn.env_origins_grid(N, cfg.env_spacing).to(dev)
    origins = origins_all[sl].contiguous()
    keys = ("curr_target_index", "swing_leg", "target_reach_count", "episode_length_buf", "potentials")

    def make(n, off, o, state):
        m = AllstepsMDP(n, device=dev, seed=seed, env_id_offset=off)
        m.generate_stones(o)
        m.import_state(state)
        return m

    shard_state = {k: st0[k][sl] for k in keys}
    peer = make(per, rank * per, origins, shard_state)
    nccl = make(per, rank * per, origins, shard_state)
    peer.connect_peers()
    single = make(N, 0, origins_all, {k: st0[k] for k in keys})  # every rank steps the whole thing: no broadcast needed
    o_p, o_n, o_s = StepBuffers(per, dev), StepBuffers(per, dev), StepBuffers(N, dev)
    gbuf = torch.zeros_like(nccl.exchange_tensor)
    gen = torch.Generator(device=dev).manual_seed(seed)  # same sequence on every rank
    ok_pn = ok_single = True
    level0, promotions = 0, 0
    for _ in range(steps):
        st = single.export_state()
        d = syn.random_physics_state(cfg, st["steps_pos"], st["curr_target_index"], st["swing_leg"], gen,
                                     fall_fraction=0.05)
        d.pop("root_ang_vel_w", None)
        ds = {k: v[sl].contiguous() for k, v in d.items()}
        v_all, v_sh = PhysicsViews.from_dict(d, origins_all), PhysicsViews.from_dict(ds, origins)
        single.step(v_all, d["actions"], o_s)
        peer.step(v_sh, ds["actions"], o_p)
        nccl.step(v_sh, ds["actions"], o_n, finish=False)
        nccl.fold_stats()
        gbuf.copy_(nccl.exchange_tensor)
        nccl.finish_step(nccl.all_reduce_exchange(gbuf, dist))
        torch.cuda.synchronize(dev)
        for name in ("obs", "reward", "terminated", "time_out"):
            a, b, c = getattr(o_p, name), getattr(o_n, name), getattr(o_s, name)[sl]
            ok_pn &= bool(torch.equal(a, b))
            ok_single &= bool(torch.equal(a, c))
        sa, sb, sc_ = peer.export_state(), nccl.export_state(), single.export_state()
        for k in ("curr_target_index", "swing_leg", "target_reach_count", "curriculum", "potentials"):
            ok_pn &= bool(torch.equal(sa[k], sb[k]))
            ok_single &= bool(torch.equal(sa[k], sc_[k][sl]))
        lvl = int(sc_["curriculum"].max())
        promotions += int(lvl != level0)
        level0 = lvl
    flags = torch.tensor([int(ok_pn), int(ok_single), int(peer.peer_status()["timeouts"] == 0)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res = {"peer_eq_nccl": bool(flags[0]), "shards_eq_single_handle": bool(flags[1]),
           "no_peer_timeouts": bool(flags[2]), "promotions": promotions, "global_envs": N, "steps": steps}
    res["ok"] = res["peer_eq_nccl"] and res["shards_eq_single_handle"] and res["no_peer_timeouts"] and promotions > 0
    del peer, nccl, single
    return res


def main_b200(args):
    import torch

    from allsteps_isaaclab_b200 import build as _build
    _build.build()  # in-tree library (prebuilt .so travels with the repo; rebuilt only if stale)
    from allsteps_isaaclab_b200 import synthetic as syn
    from allsteps_isaaclab_b200.config import AllstepsCfg
    from allsteps_isaaclab_b200.mdp import AllstepsMDP, StepBuffers
    from allsteps_isaaclab_b200.sharding import StatsReducer
    from allsteps_isaaclab_b200.workload import ChainedWorkload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the Allsteps MDP step has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    cfg = AllstepsCfg()
    N = args.envs_per_gpu

    def make(num_envs, seed=1234, period=None, fall_fraction=None, intended_regen=None, grid_bins=None,
             layout="dense", env_id_offset=None, peers=False, close_step=None):
        fall_fraction = args.fall_fraction if fall_fraction is None else fall_fraction
        intended_regen = args.intended_regen if intended_regen is None else intended_regen
        grid_bins = args.grid_bins if grid_bins is None else grid_bins
        origins = syn.env_origins_grid(num_envs, cfg.env_spacing).to(dev)
        mdp = AllstepsMDP(num_envs, device=dev, seed=seed,
                          env_id_offset=rank * num_envs if env_id_offset is None else env_id_offset,
                          intended_regen=intended_regen, grid_bins=grid_bins)
        mdp.generate_stones(origins)
        st0 = syn.random_mdp_state(cfg, num_envs, torch.Generator().manual_seed(seed + rank))
        mdp.import_state({k: st0[k] for k in ("curr_target_index", "swing_leg", "target_reach_count",
                                              "episode_length_buf", "potentials")})
        if peers:
            mdp.connect_peers()
        out = StepBuffers(num_envs, dev)
        wl = ChainedWorkload(mdp, origins, out, period or args.input_sets, seed + rank, cfg, fall_fraction, layout,
                             stones_change=bool(intended_regen or grid_bins), close_step=close_step)
        return wl, origins

    def reduce_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    side = torch.cuda.Stream(dev) if dist is not None else None
    reducer = StatsReducer(dev) if dist is not None else None
    gbuf = [None]  # a whole AsExchange record (statistics + grid outcomes)

    def nccl_close(m, v, d, o):  # config-4 semantics through the plain library route (--global-promotion nccl)
        if gbuf[0] is None:
            gbuf[0] = torch.zeros_like(m.exchange_tensor)
        m.step(v, d["actions"], o, finish=False)
        m.fold_stats()
        gbuf[0].copy_(m.exchange_tensor)
        m.finish_step(m.all_reduce_exchange(gbuf[0], dist))

    if args.global_promotion and dist is not None:  # the headline workload itself with a global promotion route
        wl, origins = make(N, peers=args.global_promotion == "peer",
                           close_step=nccl_close if args.global_promotion == "nccl" else None)
    else:
        wl, origins = make(N)
    mdp, out = wl.mdp, wl.out
    wl_rates = (wl.advance_rate, wl.reset_rate, wl.set_bytes, wl.rewind_bytes)

    def stats_hook(i):
        if dist is not None and args.stats_interval and not args.global_promotion and (i + 1) % args.stats_interval == 0:
            # episode / curriculum statistics: summed over ranks off the step path (side stream, NCCL)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                reducer.start(mdp.stats_tensor)

    with ClockSampler(local_rank) as clocks:
        ms_total, launches = time_cycle(torch, wl, args.steps, args.warmup, dist, stats_hook)
        # the dominant kernel alone (rank-local, timed by events on its stream), same clock record
        k_avg, k_med = time_kernel_only(torch, wl, min(args.steps, 200))
        step_us = time_per_step(torch, wl, min(max(args.steps, 50), 200))
    ms_total = reduce_max(ms_total)
    ms_step = ms_total / args.steps
    value = N * world * args.steps / (ms_total * 1e-3)
    if reducer is not None and reducer.work is not None:
        reducer.wait()
    stats = mdp.read_stats()

    peak, peak_src = measured_peaks()
    # k_step<fused> moves everything except the 24 B/env of contact vectors, which k_prepare* fetches for it
    b_kernel = B_ALG - 24
    achieved_kernel = N * b_kernel / (k_avg * 1e-3) / 1e9
    achieved_step = N * B_ALG / (ms_step * 1e-3) / 1e9
    traffic = profiled_traffic()

    e2e = None
    if not args.no_e2e:
        e_steps = args.e2e_steps or min(args.steps, 24)
        e_steps += e_steps % 2
        variants = {}
        for name, zc in (("copy_all_inputs", False), ("zero_copy_contact_matrices", True)):
            e_ms, h2d_b, d2h_b, _ = time_e2e(torch, wl, origins, e_steps, 4, zero_copy_contact=zc)
            e_ms = reduce_max(e_ms)
            variants[name] = {"value": N * world * e_steps / (e_ms * 1e-3), "h2d_bytes_per_step": h2d_b,
                              "d2h_bytes_per_step": d2h_b, "ms_per_step": e_ms / e_steps,
                              "host_link_GBps_per_gpu": (h2d_b + d2h_b) / (e_ms / e_steps * 1e-3) / 1e9}
        best = max(variants, key=lambda k: variants[k]["value"])
        e2e = {"value": variants[best]["value"], "unit": UNIT,
               "h2d_bytes_per_step": variants[best]["h2d_bytes_per_step"],
               "d2h_bytes_per_step": variants[best]["d2h_bytes_per_step"], "steps": e_steps,
               "ms_per_step": variants[best]["ms_per_step"], "mode": best, "variants": variants,
               "bound": "host<->device copies (PCIe; at N > 1 all ranks share the host's memory system / one NUMA "
                        f"node): {variants[best]['host_link_GBps_per_gpu'] * world:.0f} GB/s aggregate over "
                        f"{world} GPU(s).  Under Isaac Lab the state is device resident; this leg is synthetic.",
               "note": "pinned host buffers; H2D of step t+1 and D2H of step t-1 overlap the kernel of step t; in "
                       "zero_copy_contact_matrices the (N,1,20,3) contact tensors stay in pinned host memory and "
                       "k_prepare_paired reads the current stone's vectors through PCIe"}

    blocks = {}
    extra = not args.no_extra_blocks
    if extra and world > 1:
        # ---- BASELINE config 4: 1,048,576 envs SPLIT over the ranks, promotion on the global mean (peer memory)
        del wl
        torch.cuda.empty_cache()
        per = C4_GLOBAL_ENVS // world
        w4, _ = make(per, seed=4321, peers=True)
        ms4, l4 = time_cycle(torch, w4, max(args.steps, 200), args.warmup, dist)
        ms4 = reduce_max(ms4) / max(args.steps, 200)
        status = w4.mdp.peer_status()
        blocks["c4"] = {"value": C4_GLOBAL_ENVS / (ms4 * 1e-3), "unit": UNIT, "us_per_step": ms4 * 1e3,
                        "global_envs": C4_GLOBAL_ENVS, "envs_per_gpu": per, "scaling": "strong",
                        "promotion": "global mean; step counters summed by one kernel over NVLink peer memory every step",
                        "efficiency_vs_one_gpu_at_1M": (C4_GLOBAL_ENVS / (ms4 * 1e-3)) / value,
                        "peer_exchange_timeouts": status["timeouts"], "kernels_per_step": l4 / max(args.steps, 200),
                        "advance_rate": w4.advance_rate, "reset_rate": w4.reset_rate}
        del w4
        blocks["c4"]["selfcheck"] = selfcheck_global_promotion(torch, dist, dev, rank, world)

    if extra and rank == 0 and world == 1:
        del wl
        torch.cuda.empty_cache()
        k_small = max(args.steps, 500)

        def small_block(n, **kw):
            w, w_org = make(n, seed=99, **kw)
            ms2, l2 = time_cycle(torch, w, k_small, args.warmup)
            ms3, l3 = time_cycle_graph(torch, w, k_small, args.warmup)
            best = min(ms2, ms3)
            return {"us_per_step": 1e3 * best / k_small, "value": n * k_small / (best * 1e-3), "unit": UNIT,
                    "envs": n, "kernels_per_step": l3 / k_small,
                    "mode": "library calls" if ms2 <= ms3 else "one CUDA-graph replay per step",
                    "us_per_step_library_calls": 1e3 * ms2 / k_small, "us_per_step_graph_replay": 1e3 * ms3 / k_small,
                    "advance_rate": w.advance_rate, "reset_rate": w.reset_rate,
                    "bound": "launch latency (working set is L2 resident)"}, w, w_org

        # ---- BASELINE config 2: 4096 envs, rl_games' default batch scale
        blocks["c2_4096"], w2, o2 = small_block(4096)
        ms_t, l_t = time_three_call(torch, w2, o2, k_small, args.warmup)
        ms_d, l_d = time_three_call(torch, w2, o2, k_small, args.warmup, device_reset_list=True)
        three = {"4096": {"us_per_step": 1e3 * ms_t / k_small, "kernels_per_step": l_t / k_small,
                          "fused_us_per_step": blocks["c2_4096"]["us_per_step_library_calls"],
                          "device_list_us_per_step": 1e3 * ms_d / k_small, "device_list_kernels_per_step": l_d / k_small}}
        del w2
        # ---- BASELINE config 3: 65536 envs with the pitch x yaw grid curriculum (extension)
        blocks["c3_65536_grid"], w3, _ = small_block(65536, grid_bins=11)
        blocks["c3_65536_grid"]["bound"] = "launch latency (4 dependent launches: k_prepare, k_step, k_reset_rows with the grid turnover, k_fixup_finish; working set is L2 resident)"
        del w3
        w65, o65 = make(65536, seed=99)
        ms_f, _ = time_cycle(torch, w65, k_small, args.warmup)
        ms_t, l_t = time_three_call(torch, w65, o65, k_small, args.warmup)
        ms_d, l_d = time_three_call(torch, w65, o65, k_small, args.warmup, device_reset_list=True)
        three["65536"] = {"us_per_step": 1e3 * ms_t / k_small, "kernels_per_step": l_t / k_small,
                          "fused_us_per_step": 1e3 * ms_f / k_small,
                          "device_list_us_per_step": 1e3 * ms_d / k_small, "device_list_kernels_per_step": l_d / k_small}
        del w65
        # ---- the 3-call path at the headline size
        k_big = min(max(args.steps, 40), 200)
        w1m, o1m = make(N, seed=77)
        ms_f, _ = time_cycle(torch, w1m, k_big, args.warmup)
        ms_t, l_t = time_three_call(torch, w1m, o1m, k_big, args.warmup)
        ms_d, l_d = time_three_call(torch, w1m, o1m, k_big, args.warmup, device_reset_list=True)
        three[str(N)] = {"us_per_step": 1e3 * ms_t / k_big, "kernels_per_step": l_t / k_big,
                         "fused_us_per_step": 1e3 * ms_f / k_big, "ratio_to_fused": ms_t / ms_f,
                         "device_list_us_per_step": 1e3 * ms_d / k_big, "device_list_kernels_per_step": l_d / k_big,
                         "device_list_ratio_to_fused": ms_d / ms_f}
        three["what"] = ("us_per_step: as_step_pass1 -> host .nonzero() of reset_buf (DRL:359, a device->host sync, "
                         "which Isaac Lab's DirectRLEnv.step does itself) -> as_reset(ids) -> as_step_pass2: the path the "
                         "DirectRLEnv hooks take when PhysX sits between the writes of ENV:563-565 and pass 2.  "
                         "device_list_us_per_step: the same three calls with as_reset taking the envs pass 1 flagged "
                         "from the id list it compacted on the device -- no host round trip")
        blocks["three_call"] = three
        del w1m
        torch.cuda.empty_cache()
        # ---- BASELINE config 5: reset-heavy stress with stone regeneration
        w5, _ = make(N, seed=55, period=4, fall_fraction=0.3, intended_regen=True)
        ms5, l5 = time_cycle(torch, w5, k_big, args.warmup)
        b5 = B_ALG + w5.reset_rate * 244 + (w5.mdp.read_stats()["n_regenerated"] / N) * 320
        blocks["c5_reset_heavy"] = {"us_per_step": 1e3 * ms5 / k_big, "value": N * k_big / (ms5 * 1e-3), "unit": UNIT,
                                    "envs": N, "reset_rate": w5.reset_rate, "advance_rate": w5.advance_rate,
                                    "regenerated_per_step": w5.mdp.read_stats()["n_regenerated"],
                                    "kernels_per_step": l5 / k_big, "algorithmic_bytes_per_env_step": b5,
                                    "roofline_frac": N * b5 / (ms5 / k_big * 1e-3) / 1e9 / peak,
                                    "state_rewind": "every 4 steps incl. the 320 B/env stone rows"}
        del w5
        torch.cuda.empty_cache()
        # ---- the layout Isaac Lab hands out: (N,13) root_state_w slices, (N,17,13) body_state_w slice
        wi, _ = make(N, seed=66, period=4, layout="isaac")
        msi, li = time_cycle(torch, wi, k_big, args.warmup)
        blocks["isaac_layout"] = {"us_per_step": 1e3 * msi / k_big, "value": N * k_big / (msi * 1e-3), "unit": UNIT,
                                  "envs": N, "kernels_per_step": li / k_big,
                                  "roofline_frac": N * B_ALG / (msi / k_big * 1e-3) / 1e9 / peak,
                                  "what": "root pos/quat/lin vel as slices of one (N,13) root_state_w tensor, body "
                                          "positions as a slice of the (N,17,13) body_state_w tensor "
                                          "(articulation_data.py:366-380,430-449)"}
        del wi
        torch.cuda.empty_cache()

    if extra and rank == 0 and world == 1:
        # ---- mirror-symmetry augmentation of a PPO rollout (horizon 32)
        m_f1 = AllstepsMDP(4096, device=dev, seed=3)
        blocks["f1_mirror"] = {"32x4096": time_mirror(torch, m_f1, m_f1.cfg, 32 * 4096, 200, peak),
                               "32x65536": time_mirror(torch, m_f1, m_f1.cfg, 32 * 65536, 40, peak),
                               "what": "vstack((x, mirrored(x))) of obses (rows,59), actions and mus (rows,21) of "
                                       "A2CAgentSymmetry.play_steps: 1212 algorithmic bytes per row (read once, "
                                       "written twice), output buffers allocated inside the timed call like the "
                                       "reference's"}
        del m_f1
        torch.cuda.empty_cache()
        # ---- the action path
        blocks["f2_action"] = {"1048576": time_action(torch, dev, N, 120, peak),
                               "4096": time_action(torch, dev, 4096, 300, peak),
                               "what": "as_apply_action per call (Isaac Lab makes `decimation` = 4 calls per env step); "
                                       "4096 envs: the host's launch path, not the kernel"}
        torch.cuda.empty_cache()

    if extra and rank == 0 and world == 1:
        # ---- the reference's algorithm as eager torch on this GPU: what an Isaac Lab user gets today
        try:
            eager = {str(n): run_eager_cuda_port(torch, dev, n, k, 3) for n, k in ((4096, 30), (N, 8))}
            for v in eager.values():
                v["unit"] = UNIT
            eager["what"] = ("the oracle port (the reference's ops, bit-identical to it on CPU) with CUDA tensors: eager "
                             "torch kernels + the reference's own host syncs; wall time per step")
            blocks["eager_torch_cuda"] = eager
        except Exception as exc:  # a baseline, not the product: never fail the bench on it
            blocks["eager_torch_cuda"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N=1 only (rank 0 host cores)
        r = run_cpu_port(CPU_SAMPLE_ENVS, 24, 3)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"{CPU_SAMPLE_ENVS} envs x 24 steps of the same synthetic workload "
                         f"({r['ms_per_step']:.1f} ms/step), torch {torch.__version__} CPU; the port omits the "
                         "reference's six debug clones (ENV:257-266)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Allsteps-v0 fused MDP step, {N} envs per GPU, "
                                   f"{100.0 * wl_rates[1]:.1f}% of envs resetting and {100.0 * wl_rates[0]:.1f}% "
                                   "advancing a stone per step"
                                   + (", stone regeneration on reset" if args.intended_regen else "")
                                   + (f", {args.grid_bins}x{args.grid_bins} grid curriculum" if args.grid_bins else ""),
                       "envs_per_gpu": N, "global_envs": N * world, "parallelism": f"env-id shards x{world}",
                       "advance_rate": wl_rates[0], "reset_rate": wl_rates[1],
                       "l2_policy": f"cycle of {args.input_sets} input sets of {wl_rates[2] / 1e6:.0f} MB each "
                                    "(larger than the 126 MB L2), each generated from the MDP state it meets",
                       "state_rewind": f"every {args.input_sets} steps, inside the timed region "
                                       f"({wl_rates[3] / 1e6:.0f} MB device-to-device)",
                       "promotion": ({"nccl": "global mean, NCCL all-reduce of the step counters every step",
                                      "peer": "global mean, step counters summed by one kernel over NVLink peer "
                                              "memory every step"}[args.global_promotion]
                                     if (args.global_promotion and world > 1)
                                     else "shard-local (reference --distributed semantics)"),
                       "stats_allreduce_interval": (args.stats_interval if world > 1 and not args.global_promotion
                                                    else 0)},
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved_step, "peak": peak, "unit": "GB/s",
                         "frac": achieved_step / peak,
                         "traffic": traffic["dram_bytes_per_step"] if traffic and "dram_bytes_per_step" in traffic
                         else None,
                         "what": "whole step (SURVEY 8d): 652 algorithmic bytes per env-step over ms_per_step; "
                                 "kernels k_prepare_paired + k_step<fused> + k_fixup_finish",
                         "algorithmic_bytes_per_env_step": B_ALG, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved_step / 8000.0,
                         "kernel": {"name": "as::k_step<fused>", "achieved": achieved_kernel,
                                    "frac": achieved_kernel / peak, "ms_avg": k_avg, "ms_median": k_med,
                                    "algorithmic_bytes_per_env_step": b_kernel,
                                    "traffic": traffic["dram_bytes_per_launch"] if traffic else None}},
            "cpu_baseline": cpu,
            "step_us": step_us,
            "clocks": clocks.summary(),
            "step_stats": {k: stats[k] for k in ("n_reset", "n_terminated", "n_time_out", "n_advanced", "level")},
            **blocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
