"""Host-side cost of the DirectRLEnv hooks (B1 face): wall clock per env step at small batches, where the GPU work is a
few tens of microseconds and the Python glue decides the rate.  `--profile` prints the cProfile top of the loop."""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from scenario import Scenario
import test_gpu_faces as tf
from allsteps_isaaclab_b200.env import StandaloneAllstepsEnv


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
    sc = Scenario(N, seed=17, full_bodies=True)
    st0 = sc.initial_mdp_state()
    orc = tf._oracle(sc, st0)
    phys = sc.physics(orc.steps_pos, orc.curr_target_index, orc.swing_leg)
    robot, left, right = tf._fake_world(sc, phys)
    scene = types.SimpleNamespace(env_origins=sc.env_origins.cuda())
    env = StandaloneAllstepsEnv(robot, left, right, scene, "cuda:0", sc.cfg, seed=17)
    env.reset()
    actions = phys["actions"].cuda()

    def loop(n):
        for _ in range(n):
            env._pre_physics_step(actions)
            for _ in range(4):
                env._apply_action()
            env.post_physics_step(actions)

    from allsteps_isaaclab_b200.mdp import AllstepsMDP, PhysicsViews
    fast_cached, fast_stream = PhysicsViews.cached, AllstepsMDP._stream

    def slow_cached(holder, tensors, body_rows, quat_xyzw=False):  # what the hooks did before: rebuild every time
        return PhysicsViews(**dict(zip(PhysicsViews._FIELDS, tensors)), body_rows=body_rows, quat_xyzw=quat_xyzw)

    def slow_stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    for rep in range(2):
        for name, c, st in (("rebuilt views + Stream object", slow_cached, slow_stream),
                            ("cached views + raw stream handle", fast_cached, fast_stream)):
            PhysicsViews.cached = staticmethod(c) if c is slow_cached else fast_cached
            AllstepsMDP._stream = st
            loop(20)
            torch.cuda.synchronize()
            n = 300
            t0 = time.perf_counter()
            loop(n)
            t_issue = time.perf_counter() - t0
            torch.cuda.synchronize()
            t_all = time.perf_counter() - t0
            print(f"N={N} {name}: {1e6 * t_all / n:.1f} us per env step wall clock "
                  f"({1e6 * t_issue / n:.1f} us host issue)")
    if "--profile" in sys.argv:
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        loop(100)
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(28)


main()
