"""Cross-shard promotion over NVLink peer memory against the NCCL route (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_check.py

Every rank owns one env-id shard of the same seeded scenario and steps it twice per step: once closing the step with
as_fold_stats -> NCCL all-reduce -> as_finish_step(global) and once with the peer-memory exchange kernel
(AllstepsMDP.connect_peers).  Outputs, MDP state and levels must be bit-identical between the two on every rank, the
global counters must equal the NCCL sum, and promotions must actually occur.  Then both routes are timed.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

from allsteps_isaaclab_b200.mdp import AllstepsMDP, PhysicsViews, StepBuffers
from scenario import Scenario


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    per = int(os.environ.get("PEER_CHECK_ENVS", "4096"))
    steps = int(os.environ.get("PEER_CHECK_STEPS", "24"))
    N, seed = per * world, 17
    sc = Scenario(N, seed=seed)                     # same seed on every rank: identical CPU-side scenario
    st0 = sc.initial_mdp_state()
    st0["curr_target_index"] = torch.randint(11, 20, (N,), generator=sc.gen)   # near the promotion threshold
    sl = slice(rank * per, (rank + 1) * per)
    origins = sc.env_origins[sl].to(dev)
    mdps = [AllstepsMDP(per, device=dev, seed=seed, env_id_offset=rank * per) for _ in range(2)]
    for m in mdps:
        m.generate_stones(origins)
        m.import_state({k: st0[k][sl] for k in ("curr_target_index", "swing_leg", "target_reach_count",
                                                "episode_length_buf", "potentials")})
    nccl_mdp, peer_mdp = mdps
    assert peer_mdp.connect_peers() == (world, rank)
    outs = [StepBuffers(per, dev), StepBuffers(per, dev)]
    g = torch.zeros_like(nccl_mdp.stats_tensor)
    promotions = 0
    level_before = 0
    for step in range(steps):
        st = nccl_mdp.export_state()
        # the scenario needs the stones / indices of ALL envs: gather the shards
        full = {}
        for k in ("steps_pos", "curr_target_index", "swing_leg"):
            parts = [torch.empty_like(st[k]) for _ in range(world)]
            dist.all_gather(parts, st[k].contiguous())
            full[k] = torch.cat(parts).cpu()
        phys = sc.physics(full["steps_pos"], full["curr_target_index"], full["swing_leg"])
        d = {k: v[sl].to(dev) for k, v in phys.items()}
        views = PhysicsViews.from_dict(d, origins, sc.body_indices)
        nccl_mdp.step(views, d["actions"], outs[0], finish=False)
        nccl_mdp.fold_stats()
        g.copy_(nccl_mdp.stats_tensor)
        dist.all_reduce(g[:10])
        nccl_mdp.finish_step(g)
        peer_mdp.step(views, d["actions"], outs[1])
        torch.cuda.synchronize()
        for name in ("obs", "reward", "terminated", "time_out", "dones", "reset_joint_pos"):
            a, b = getattr(outs[0], name), getattr(outs[1], name)
            assert torch.equal(a, b), f"rank {rank} step {step}: {name} differs between the NCCL and the peer route"
        sa, sb = nccl_mdp.export_state(), peer_mdp.export_state()
        for k in sa:
            assert torch.equal(sa[k], sb[k]), f"rank {rank} step {step}: state {k}"
        assert torch.equal(peer_mdp.global_stats_tensor[:10], g[:10]), (
            f"rank {rank} step {step}: global counters {peer_mdp.global_stats_tensor[:10].tolist()} vs NCCL "
            f"{g[:10].tolist()}")
        level = int(sb["curriculum"].max())
        promotions += int(level != level_before)
        level_before = level
    status = peer_mdp.peer_status()
    assert status["timeouts"] == 0 and status["world"] == world, status
    assert promotions > 0, "no promotion happened; the check would prove nothing"

    # ---- timing of the three ways to close a step (device time, max over ranks)
    def timed(fn, reps=200):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    local_mdp = AllstepsMDP(per, device=dev, seed=seed, env_id_offset=rank * per)
    local_mdp.generate_stones(origins)
    out_l = StepBuffers(per, dev)

    def via_nccl():
        nccl_mdp.step(views, d["actions"], outs[0], finish=False)
        nccl_mdp.fold_stats()
        g.copy_(nccl_mdp.stats_tensor)
        dist.all_reduce(g[:10])
        nccl_mdp.finish_step(g)

    t_local = timed(lambda: local_mdp.step(views, d["actions"], out_l))
    t_nccl = timed(via_nccl)
    t_peer = timed(lambda: peer_mdp.step(views, d["actions"], outs[1]))
    assert peer_mdp.peer_status()["timeouts"] == 0
    if rank == 0:
        print(f"peer_check OK: world {world}, {per} envs per rank, {steps} steps, {promotions} promotions; "
              f"us/step: shard-local {t_local:.1f}, NCCL all-reduce {t_nccl:.1f}, peer-memory exchange {t_peer:.1f}",
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
