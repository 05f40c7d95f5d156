#!/usr/bin/env python
"""bench.py -- Allsteps-v0 batched MDP step on B200 (and the reference's CPU implementation beside it).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path, N ranks via torchrun
    python bench.py --impl reference --steps K --warmup W      # the reference algorithm on the host cores

A "step" is one pass of the hot path (direct_rl_env.py:351-375 of the reference: episode counter, pass 1, dones,
rewards, masked reset incl. the PhysX start-pose rows, pass 2, observations) over one batch of synthetic
articulation state.  Workload at every N: 1,048,576 envs per GPU (weak scaling; envs are
sharded by env id, no collective on the step path; step statistics are all-reduced over NCCL off the step path).

JSON keys: `value` = env-steps/s with inputs resident in HBM (CUDA events, max over ranks); `e2e` = the same metric
through the public API with HOST buffers (pinned host -> device copy of every step's inputs and device -> host
read of its results inside the timed region); `roofline` = algorithmic bytes (652 B per env-step)
over the live-measured duration of the fused step kernel against the measured HBM copy peak; `cpu_baseline` = the
CPU oracle port on a bounded sample.
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
METRIC = "MDP env-steps/sec"
UNIT = "env-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--input-sets", type=int, default=4, help="distinct synthetic states rotated through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 24)")
    ap.add_argument("--stats-interval", type=int, default=16, help="all-reduce step statistics every k steps")
    ap.add_argument("--global-promotion", nargs="?", const="peer", default=None, choices=["peer", "nccl"],
                    help="N > 1: apply the promotion rule ENV:471 to the mean over ALL shards every step (config 4). "
                         "peer (default when given): one exchange kernel over NVLink peer memory; nccl: as_fold_stats + "
                         "NCCL all-reduce + as_finish_step(global). Without the flag promotion is shard-local.")
    ap.add_argument("--small-sizes", default="4096,65536", help="extra env counts timed for latency (N=1 only)")
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
def run_cpu_port(num_envs: int, steps: int, warmup: int, seed: int = 1234):
    """Times the CPU oracle port (the reference's algorithm, op for op in torch) on all host cores."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from scenario import Scenario, install_mdp_state
    from oracle import allsteps_oracle as ao

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc = Scenario(num_envs, seed=seed)
    orc = ao.AllstepsOracle(sc.cfg, num_envs, sc.env_origins, sc.joint_limits, sc.body_indices, sc.stone_uniforms(0))
    install_mdp_state(orc, sc.initial_mdp_state())
    pool = [sc.physics(orc.steps_pos, orc.curr_target_index, orc.swing_leg) for _ in range(4)]
    mirror_u, noise_u = sc.reset_uniforms(0)
    times = []
    for i in range(warmup + steps):
        phys = pool[i % len(pool)]
        t0 = time.perf_counter()
        orc.step(phys, phys["actions"], mirror_u, noise_u, None)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    return {"value": num_envs * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores,
            "threads": torch.get_num_threads()}


def main_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (it is pure Python/torch and cannot be
    installed on the GPU box, so the bit-identical oracle port stands in -- kind "port"), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample per step: ~1.1 us of host time per env-step => keep K steps within about a minute
    envs = CPU_SAMPLE_ENVS
    while envs > 4096 and envs * args.steps * 1.1e-6 > 60.0:
        envs //= 2
    r = run_cpu_port(envs, args.steps, max(args.warmup, 3))
    sample = f"{envs} envs per step (1/{ENVS_PER_GPU // envs} of the 1,048,576-env workload), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Allsteps-v0 fused MDP step, 1,048,576 envs per GPU (reference arm: bounded "
                               "sample per step on host cores)", "envs_per_step": envs},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- CUDA arm
def build_pool(torch, syn, cfg, mdp, origins, sets, device, seed, fall_fraction=0.02):
    """`sets` distinct synthetic post-physics states, generated on the device (throughput only)."""
    from allsteps_isaaclab_b200.mdp import PhysicsViews

    gen = torch.Generator(device=device).manual_seed(seed)
    st = mdp.export_state()
    pool = []
    for _ in range(sets):
        d = syn.random_physics_state(cfg, st["steps_pos"], st["curr_target_index"], st["swing_leg"], gen,
                                     fall_fraction=fall_fraction)
        d.pop("root_ang_vel_w", None)
        pool.append((PhysicsViews.from_dict(d, origins), d))
    return pool


def time_steps(torch, mdp, pool, out, steps, warmup, dist=None, stats_interval=0, stats_buf=None, side=None,
               global_promotion=False):
    dev = mdp.device

    def one_step(v, d):
        if global_promotion == "nccl" and dist is not None:
            # the one cross-env dependency of the path, ENV:471 -- the additive counters of all shards
            mdp.step(v, d["actions"], out, finish=False)
            mdp.fold_stats()
            stats_buf.copy_(mdp.stats_tensor)      # a whole AsStats; its first 10 int64 are the additive counters
            dist.all_reduce(stats_buf[:10])
            mdp.finish_step(stats_buf)
        else:  # shard-local, or peers connected: the exchange kernel is part of the step
            mdp.step(v, d["actions"], out)

    for i in range(warmup):
        v, d = pool[i % len(pool)]
        one_step(v, d)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize(dev)
    launches0 = mdp.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        v, d = pool[i % len(pool)]
        one_step(v, d)
        if dist is not None and stats_interval and not global_promotion and (i + 1) % stats_interval == 0:
            # episode / curriculum statistics: summed over ranks off the step path (side stream, NCCL)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                stats_buf.copy_(mdp.stats_tensor)
                dist.all_reduce(stats_buf[:10])
    e1.record()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1), mdp.launch_count - launches0


def time_steps_graph(torch, mdp, pool, out, steps, warmup):
    """Same loop with every input set's step captured once as a CUDA graph and replayed."""
    dev = mdp.device
    graphs = [mdp.capture_step(v, d["actions"], out) for v, d in pool]
    for i in range(warmup):
        graphs[i % len(graphs)].replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        graphs[i % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1), steps * graphs[0].kernels_per_replay


def time_kernel_only(torch, mdp, pool, out, steps):
    """Average duration of the dominant kernel alone, k_step<fused>: the library records a pair of CUDA events on the
    launching stream immediately around that launch (as_set_timing_events).  Also returns the duration of the whole
    device-side step (contact-gather kernel + step kernel + finish kernel) from events around the two API calls."""
    from allsteps_isaaclab_b200 import _cabi

    dev = mdp.device
    pairs = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b.record()  # creates the underlying cudaEvent_t
        pairs.append((a, b))
    torch.cuda.synchronize(dev)
    for i in range(steps):
        v, d = pool[i % len(pool)]
        a, b = pairs[i]
        _cabi.check(mdp.lib.as_set_timing_events(mdp.handle, a.cuda_event, b.cuda_event), "as_set_timing_events")
        mdp.step(v, d["actions"], out)
    _cabi.check(mdp.lib.as_set_timing_events(mdp.handle, None, None), "as_set_timing_events")
    torch.cuda.synchronize(dev)
    ms = sorted(a.elapsed_time(b) for a, b in pairs)
    return sum(ms) / len(ms), ms[len(ms) // 2]


def time_e2e(torch, mdp, pool, origins, out, steps, warmup, zero_copy_contact=False):
    """Same metric through the public API with HOST buffers: every step copies that step's inputs from pinned host
    memory, runs the fused step, and reads the results back to pinned host memory.  Copies of step t+1 / t-1
    overlap the kernel of step t on separate streams (double-buffered device inputs and outputs).

    zero_copy_contact: the two (N,1,20,3) contact matrices (59 % of the input bytes, of which the step needs 24 B per
    env) are NOT copied; the C ABI is handed the pinned host tensors themselves (device-accessible under unified
    addressing) and the contact-gather kernel fetches just the current stone's vectors across PCIe."""
    from allsteps_isaaclab_b200.mdp import PhysicsViews, StepBuffers

    dev = mdp.device
    N = mdp.num_envs
    keys = ["root_pos_w", "root_quat_w", "root_lin_vel_w", "body_pos_w", "joint_pos", "joint_vel",
            "force_matrix_right", "force_matrix_left", "actions"]
    host_sets = []
    for _, d in pool[:2]:
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
    outs = [out, StepBuffers(N, dev)]
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


def main_b200(args):
    import torch

    from allsteps_isaaclab_b200 import build as _build
    _build.build()  # in-tree library (prebuilt .so travels with the repo; rebuilt only if stale)
    from allsteps_isaaclab_b200 import synthetic as syn
    from allsteps_isaaclab_b200.config import AllstepsCfg
    from allsteps_isaaclab_b200.mdp import AllstepsMDP, StepBuffers

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

    def make(num_envs, seed=1234):
        origins = syn.env_origins_grid(num_envs, cfg.env_spacing).to(dev)
        mdp = AllstepsMDP(num_envs, device=dev, seed=seed, env_id_offset=rank * num_envs,
                          intended_regen=args.intended_regen, grid_bins=args.grid_bins)
        mdp.generate_stones(origins)
        st0 = syn.random_mdp_state(cfg, num_envs, torch.Generator().manual_seed(seed + rank))
        mdp.import_state({k: st0[k] for k in ("curr_target_index", "swing_leg", "target_reach_count",
                                              "episode_length_buf", "potentials")})
        pool = build_pool(torch, syn, cfg, mdp, origins, args.input_sets, dev, seed + rank, args.fall_fraction)
        return mdp, origins, pool, StepBuffers(num_envs, dev)

    mdp, origins, pool, out = make(N)
    if args.global_promotion == "peer" and dist is not None:
        mdp.connect_peers()
    side = torch.cuda.Stream(dev) if dist is not None else None
    stats_buf = torch.zeros_like(mdp.stats_tensor) if dist is not None else None

    with ClockSampler(local_rank) as clocks:
        ms_total, launches = time_steps(torch, mdp, pool, out, args.steps, args.warmup, dist,
                                        args.stats_interval, stats_buf, side, args.global_promotion)
        # roofline of the dominant kernel (rank-local, timed alone on its stream), same clock record
        k_avg, k_med = time_kernel_only(torch, mdp, pool, out, min(args.steps, 200))
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = N * world * args.steps / (ms_total * 1e-3)
    stats = mdp.read_stats()

    peak, peak_src = measured_peaks()
    # k_step<fused> moves everything except the 24 B/env of contact vectors, which k_contact_gather fetches for it
    b_kernel = B_ALG - 24
    achieved = N * b_kernel / (k_avg * 1e-3) / 1e9
    achieved_step = N * B_ALG / (ms_step * 1e-3) / 1e9
    traffic = profiled_traffic()

    e2e = None
    if not args.no_e2e:
        e_steps = args.e2e_steps or min(args.steps, 24)
        variants = {}
        for name, zc in (("copy_all_inputs", False), ("zero_copy_contact_matrices", True)):
            e_ms, h2d_b, d2h_b, _ = time_e2e(torch, mdp, pool, origins, out, e_steps, 3, zero_copy_contact=zc)
            te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            variants[name] = {"value": N * world * e_steps / (float(te.item()) * 1e-3), "h2d_bytes_per_step": h2d_b,
                              "d2h_bytes_per_step": d2h_b, "ms_per_step": float(te.item()) / e_steps}
        best = max(variants, key=lambda k: variants[k]["value"])
        e2e = {"value": variants[best]["value"], "unit": UNIT,
               "h2d_bytes_per_step": variants[best]["h2d_bytes_per_step"],
               "d2h_bytes_per_step": variants[best]["d2h_bytes_per_step"], "steps": e_steps,
               "ms_per_step": variants[best]["ms_per_step"], "mode": best, "variants": variants,
               "note": "pinned host buffers; H2D of step t+1 and D2H of step t-1 overlap the kernel of step t; in "
                       "zero_copy_contact_matrices the (N,1,20,3) contact tensors stay in pinned host memory and "
                       "k_contact_gather_paired reads the current stone's vectors through PCIe"}

    small = {}
    if rank == 0 and world == 1 and args.small_sizes:
        for n in [int(x) for x in args.small_sizes.split(",") if x]:
            m2, _, p2, o2 = make(n, seed=99)
            k2 = max(args.steps, 500)
            ms2, l2 = time_steps(torch, m2, p2, o2, k2, args.warmup)
            ms3, l3 = time_steps_graph(torch, m2, p2, o2, k2, args.warmup)
            best = min(ms2, ms3)  # since the kernels are chained by programmatic dependent launch the two are close
            small[str(n)] = {"us_per_step": 1e3 * best / k2, "env_steps_per_s": n * k2 / (best * 1e-3),
                             "kernels_per_step": l3 / k2,
                             "mode": "library calls" if ms2 <= ms3 else "one CUDA-graph replay per step",
                             "us_per_step_library_calls": 1e3 * ms2 / k2, "us_per_step_graph_replay": 1e3 * ms3 / k2,
                             "bound": "launch latency (working set is L2 resident)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N=1 only (rank 0 host cores)
        r = run_cpu_port(CPU_SAMPLE_ENVS, 24, 3)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"{CPU_SAMPLE_ENVS} envs x 24 steps of the same synthetic workload "
                         f"({r['ms_per_step']:.1f} ms/step), torch {torch.__version__} CPU"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Allsteps-v0 fused MDP step, {N} envs per GPU, "
                                   f"{100.0 * stats['n_reset'] / N:.0f}% of envs resetting per step"
                                   + (", stone regeneration on reset" if args.intended_regen else "")
                                   + (f", {args.grid_bins}x{args.grid_bins} grid curriculum" if args.grid_bins else ""),
                       "envs_per_gpu": N, "global_envs": N * world, "parallelism": f"env-id shards x{world}",
                       "l2_policy": f"{args.input_sets} rotating input sets of {N * 808 / 1e6:.0f} MB each "
                                    "(larger than the 126 MB L2)",
                       "promotion": ({"nccl": "global mean, NCCL all-reduce of the step counters every step",
                                      "peer": "global mean, step counters summed by one kernel over NVLink peer "
                                              "memory every step"}[args.global_promotion]
                                     if (args.global_promotion and world > 1)
                                     else "shard-local (reference --distributed semantics)"),
                       "stats_allreduce_interval": (args.stats_interval if world > 1 and not args.global_promotion
                                                    else 0),
                       **({"peer_exchange_timeouts": mdp.peer_status()["timeouts"]}
                          if (args.global_promotion == "peer" and world > 1) else {})},
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "kernel": "as::k_step<fused>", "kernel_ms_avg": k_avg, "kernel_ms_median": k_med,
                         "algorithmic_bytes_per_env_step": b_kernel, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "whole_step": {"achieved": achieved_step, "frac": achieved_step / peak,
                                        "algorithmic_bytes_per_env_step": B_ALG,
                                        "kernels": "k_contact_gather_paired + k_step<fused> + k_fixup_finish"}},
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
            "step_stats": {k: stats[k] for k in ("n_reset", "n_terminated", "n_time_out", "n_advanced", "level")},
            "other_sizes": small,
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
