"""CPU, build container only (skipped where /root/reference is not mounted): the B2 face -- terms.py + manager_cfg.py --
driven by the reference's UNMODIFIED managers (oracle/ref_managers.py loads isaaclab/managers/*.py): the term
configurations pass `ManagerBase._resolve_common_term_cfg` (manager_base.py:219-298), the ObservationManager's
construction-time shape probe (observation_manager.py:411) steps nothing, and a replay in ManagerBasedRLEnv.step order
(manager_based_rl_env.py:203-239, `_reset_idx` :347-392) through ObservationManager / RewardManager /
TerminationManager / EventManager / CurriculumManager reproduces the reference MDP step.  Two more tests hand the order
of the calls to the reference as well: the bodies of its `ManagerBasedRLEnv.step` / `_reset_idx` drive the terms (B2), and
the bodies of its `DirectRLEnv.reset` / `step` / `_reset_idx` drive the six hooks of `env.py` (B1).

There is no GPU here, so the CUDA handle behind the terms is replaced (monkeypatched module attribute, test only) by a
stand-in with the same methods that serves the CPU oracle's numbers: what is under test is the term layer -- signatures,
parameter names, caching per `common_step_counter`, the behaviour before the first step, weights and flags of the
configuration -- against the managers that would really call it.  The same terms on the real handle are compared
with the oracle on the GPU (tests/test_gpu_faces.py)."""
from __future__ import annotations

import types

import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not mounted")


class OracleMDP:
    """Stand-in for AllstepsMDP (pass1 / reset / pass2 / generate_stones / stats_tensor) over the CPU oracle."""

    instances = []

    def __init__(self, num_envs, device=None, cfg=None, seed=0, joint_limits=None, **kw):
        from allsteps_isaaclab_b200 import synthetic as syn
        from allsteps_isaaclab_b200.config import AllstepsCfg

        self.cfg = cfg or AllstepsCfg()
        self.N, self.seed = num_envs, seed
        self.device = torch.device("cpu")
        self.joint_limits = joint_limits if joint_limits is not None else syn.joint_limits_tensor(self.cfg)
        self.launch_count = 0      # "launches": calls that would start kernels
        self.counter = 0           # the library's Philox step counter
        self.pass1_done = False
        self.stats_tensor = torch.zeros(14, dtype=torch.int64)
        self.orc = None
        OracleMDP.instances.append(self)

    def generate_stones(self, origins, env_ids=None, uniforms=None):
        from oracle import allsteps_oracle as ao
        from oracle import philox

        ids = np.arange(self.N, dtype=np.int64)
        su = torch.from_numpy(philox.stone_tables(self.seed, 0, ids, self.cfg.num_steps))
        self.orc = ao.AllstepsOracle(self.cfg, self.N, origins, self.joint_limits, self.body_rows, su)
        self.launch_count += 1

    def _load(self, views):
        t = dict(views.tensors)
        phys = {k: t[k] for k in ("root_pos_w", "root_quat_w", "root_lin_vel_w", "body_pos_w", "joint_pos", "joint_vel")}
        phys["force_matrix_left"], phys["force_matrix_right"] = t["force_matrix_left"], t["force_matrix_right"]
        phys["root_ang_vel_w"] = torch.zeros(self.N, 3)
        return phys

    def pass1(self, views, actions, buf, episode_length=None):
        o = self.orc
        o.load_physics(self._load(views))
        o.clamp_actions(actions)
        o.episode_length_buf = episode_length.clone()
        term, to = o.dones()
        o.reset_terminated[:], o.reset_time_outs[:] = term, to
        buf.reward.copy_(o.rewards())
        buf.terminated.copy_(term)
        buf.time_out.copy_(to)
        buf.dones.copy_(term | to)
        buf.obs.copy_(o.observations())
        self.stats_tensor[0] = self.N
        self.stats_tensor[8] = int(o.curr_target_index.sum())
        self.stats_tensor[10] = int(o.curriculum.max())
        self.counter += 1
        self.pass1_done = True
        self.launch_count += 1

    def reset(self, origins, env_ids, buf, episode_length=None):
        from oracle import philox

        if not self.pass1_done:  # as_reset outside a step advances the Philox step counter itself
            self.counter += 1
        self.pass1_done = False
        m, n = philox.reset_tables(self.seed, self.counter, np.arange(self.N, dtype=np.int64))
        ids = env_ids.to(torch.int64)
        self.orc.reset_rows(ids, torch.from_numpy(m)[ids], torch.from_numpy(n)[ids])
        w = self.orc.reset_writes
        k = len(ids)
        buf.reset_root_state[:k] = torch.cat((w["root_pose"], w["root_velocity"]), -1)
        buf.reset_joint_pos[:k] = w["joint_pos"]
        buf.reset_joint_vel[:k] = w["joint_vel"]
        if episode_length is not None:
            episode_length[ids] = 0
        self.launch_count += 1

    def no_reset(self):
        pass  # (the stand-in's pass 1 leaves the pass-1 observations in place)

    def apply_action(self, actions, efforts=None):
        self.orc.clamp_actions(actions)
        self.launch_count += 1
        return self.orc.joint_efforts()

    def pass2(self, views, buf):
        buf.obs.copy_(self.orc.observations())  # (the oracle's reset ended with its pass 2 on the rows it wrote)
        self.launch_count += 1


class _Scene(dict):
    env_origins = None

    def reset(self, env_ids):
        self["foot_contacts_left"].data.force_matrix_w[env_ids] = 0.0   # contact_sensor.py:155
        self["foot_contacts_right"].data.force_matrix_w[env_ids] = 0.0


def _world(sc, phys, M):
    from allsteps_isaaclab_b200.config import BODY_NAMES, JOINT_NAMES
    from oracle.ref_fake_env import FakeRobot

    world = {k: v.clone() for k, v in phys.items()}
    world["joint_pos_limits"] = sc.joint_limits.unsqueeze(0).repeat(sc.N, 1, 1)
    robot = FakeRobot(world, BODY_NAMES, JOINT_NAMES, sc.cfg.default_root_pos)
    left = types.SimpleNamespace(data=types.SimpleNamespace(force_matrix_w=world["force_matrix_left"]))
    right = types.SimpleNamespace(data=types.SimpleNamespace(force_matrix_w=world["force_matrix_right"]))
    scene = _Scene(robot=robot, foot_contacts_left=left, foot_contacts_right=right)
    scene.env_origins = sc.env_origins
    return scene, robot, left, right


def test_reference_managers_drive_the_terms(monkeypatch):
    from allsteps_isaaclab_b200 import manager_cfg, terms
    from oracle import allsteps_oracle as ao
    from oracle import ref_managers
    from scenario import Scenario, install_mdp_state

    M = ref_managers.load_managers()
    N, seed = 96, 19
    sc = Scenario(N, seed=seed, full_bodies=True)
    OracleMDP.body_rows = sc.body_indices
    OracleMDP.instances.clear()
    monkeypatch.setattr(terms, "AllstepsMDP", OracleMDP)
    st0 = sc.initial_mdp_state()
    direct = ao.AllstepsOracle(sc.cfg, N, sc.env_origins, sc.joint_limits, sc.body_indices, sc.stone_uniforms(0))
    install_mdp_state(direct, st0)
    phys = sc.physics(direct.steps_pos, direct.curr_target_index, direct.swing_leg)
    scene, robot, left, right = _world(sc, phys, M)
    dt = sc.cfg.step_dt
    env = types.SimpleNamespace(num_envs=N, device="cpu", scene=scene, common_step_counter=0, step_dt=dt,
                                max_episode_length_s=sc.cfg.episode_length_s,
                                episode_length_buf=st0["episode_length_buf"].clone(),
                                actions=torch.zeros(N, 21))

    # ---- construction: the cfgs pass the managers' validation, and the shape probe launches no MDP pass
    cfgs = manager_cfg.build_manager_cfgs(M)
    terms.binding(env, seed=seed)  # (the seed of the test's Philox tables; Isaac Lab would create it on first use)
    mdp = OracleMDP.instances[0]
    launches = mdp.launch_count
    obs_man = M.ObservationManager(cfgs["observations"], env)
    assert mdp.launch_count == launches and mdp.counter == 0, "the shape probe must not step the MDP"
    assert obs_man.group_obs_dim["policy"] == (59,)
    assert obs_man.active_terms["policy"] == [f.__name__ for f in terms.OBSERVATION_TERMS]
    rew_man = M.RewardManager(cfgs["rewards"], env)
    term_man = M.TerminationManager(cfgs["terminations"], env)
    evt_man = M.EventManager(cfgs["events"], env)
    cur_man = M.CurriculumManager(cfgs["curriculum"], env)
    assert "reset" in evt_man.available_modes
    assert rew_man.active_terms == ["allsteps"] and term_man.active_terms == ["terminated", "time_out"]
    for man in (obs_man, rew_man, term_man, evt_man, cur_man):
        assert str(man)
    # the per-term reward configuration is accepted too; a misspelt parameter is rejected by the reference's check
    M.RewardManager(manager_cfg.build_manager_cfgs(M, per_term_rewards=True)["rewards"], env)
    bad = manager_cfg.build_manager_cfgs(M)["rewards"]
    bad["allsteps"].params["robot"] = bad["allsteps"].params.pop("asset_cfg")
    with pytest.raises(ValueError, match="expects mandatory parameters"):
        M.RewardManager(bad, env)
    missing = manager_cfg.build_manager_cfgs(M, robot="no_such_robot")["terminations"]
    with pytest.raises(ValueError, match="does not exist"):
        M.TerminationManager(missing, env)
    assert mdp.launch_count == launches

    def install(phys):
        world = {k: v.clone() for k, v in phys.items()}
        robot.load_physics(world)
        left.data.force_matrix_w = world["force_matrix_left"]
        right.data.force_matrix_w = world["force_matrix_right"]
        env.actions = world["actions"]

    # ---- ManagerBasedEnv.reset() before the first step (manager_based_env.py:256-300): `_reset_idx(all ids)`, then the
    # observations, with common_step_counter still 0 -- the terms must serve what the reset event left, not step
    fresh = ao.AllstepsOracle(sc.cfg, N, sc.env_origins, sc.joint_limits, sc.body_indices, sc.stone_uniforms(0))
    phys = sc.physics(fresh.steps_pos, fresh.curr_target_index, fresh.swing_leg)
    m, n = sc.reset_uniforms(1)
    fresh.load_physics(phys)
    fresh.reset_rows(torch.arange(N), m, n)
    install(phys)
    mdp.orc.load_physics(mdp._load(terms.binding(env).views(env)))  # (the stand-in's oracle wants the tensors early)
    all_ids = torch.arange(N, dtype=torch.int64)
    cur_man.compute(env_ids=all_ids)
    scene.reset(all_ids)
    evt_man.apply(mode="reset", env_ids=all_ids, global_env_step_count=0)
    for man in (obs_man, rew_man, cur_man, evt_man, term_man):
        man.reset(all_ids)
    env.episode_length_buf[all_ids] = 0
    obs = obs_man.compute()
    assert torch.equal(obs["policy"], fresh.observations()), "observations of the initial reset"
    assert mdp.counter == 1 and not mdp.pass1_done
    assert torch.equal(robot.rec.calls["joint_state"][0], fresh.reset_writes["joint_pos"])

    # continue from a mid-episode state; the stand-in holds the same MDP state as the directly stepped oracle
    install_mdp_state(mdp.orc, st0)
    env.episode_length_buf[:] = st0["episode_length_buf"]
    assert torch.equal(mdp.orc.steps_pos, direct.steps_pos)

    n_reset = 0
    sim_step = 0
    for step in range(10):
        phys = sc.physics(direct.steps_pos, direct.curr_target_index, direct.swing_leg)
        m, n = sc.reset_uniforms(step + 2)  # initial reset: 1; pass 1 advances the counter before the reset draws
        o_obs, o_rew, o_term, o_to, o_ids = direct.step(phys, phys["actions"], m, n, None)
        install(phys)
        # ---- ManagerBasedRLEnv.step after the physics loop, manager_based_rl_env.py:203-239
        sim_step += 4
        env.episode_length_buf += 1
        env.common_step_counter += 1
        reset_buf = term_man.compute()
        assert torch.equal(term_man.terminated, o_term) and torch.equal(term_man.time_outs, o_to)
        reward = rew_man.compute(dt=dt)
        assert torch.allclose(reward, o_rew, rtol=1e-6, atol=1e-6), (reward - o_rew).abs().max()
        ids = reset_buf.nonzero(as_tuple=False).squeeze(-1)
        assert torch.equal(ids, o_ids)
        if len(ids) > 0:  # `_reset_idx`, manager_based_rl_env.py:347-392
            n_reset += len(ids)
            cur_man.compute(env_ids=ids)
            scene.reset(ids)
            evt_man.apply(mode="reset", env_ids=ids, global_env_step_count=sim_step // 4)
            log = {}
            for man in (obs_man, rew_man, cur_man, evt_man, term_man):
                log.update(man.reset(ids))
            assert 0.0 <= log["Curriculum/allsteps_level/level"] <= 9.0
            assert abs(log["Curriculum/allsteps_level/mean_target_index"]
                       - float(direct.pass1["curr_target_index"].float().mean())) < 1e-3
            env.episode_length_buf[ids] = 0
            c = robot.rec.calls
            assert torch.equal(c["root_pose"][0], direct.reset_writes["root_pose"])
            assert torch.equal(c["joint_state"][0], direct.reset_writes["joint_pos"])
            assert torch.equal(c["joint_state"][2], o_ids)
        obs = obs_man.compute()
        assert obs["policy"].shape == (N, 59)
        assert torch.equal(obs["policy"], o_obs), f"step {step}"
        assert torch.equal(env.episode_length_buf, direct.episode_length_buf)
        # one pass 1 (+ one reset + one pass 2) per env step, however many terms asked
        before = mdp.launch_count
        term_man.compute()
        obs_man.compute()
        assert mdp.launch_count == before
    assert n_reset > 0


class _Quiet:
    """Manager / simulator stand-ins that the Allsteps task does not use (actions are applied by PhysX, nothing is
    recorded, there are no commands): every method is a no-op, `reset` returns an empty log."""

    active_terms: list = []

    def __getattr__(self, name):
        if name == "reset":
            return lambda *a, **k: {}
        if name in ("has_gui", "has_rtx_sensors"):
            return lambda: False
        return lambda *a, **k: None


def test_the_reference_step_function_drives_the_terms(monkeypatch):
    """The same replay, but the ORDER of the manager calls is not restated here: the bodies of the reference's own
    `ManagerBasedRLEnv.step` and `_reset_idx` (manager_based_rl_env.py:153-239,347-392, compiled from its source by
    oracle/ref_managers.load_rl_env_methods) run on an env object that carries the reference's managers built on our
    term configuration.  Physics is the scene stand-in's `update()`: it installs the next synthetic state at the last
    of the `decimation` sub-steps."""
    from allsteps_isaaclab_b200 import manager_cfg, terms
    from oracle import allsteps_oracle as ao
    from oracle import ref_managers
    from scenario import Scenario, install_mdp_state

    M = ref_managers.load_managers()
    methods = ref_managers.load_rl_env_methods()
    N, seed = 80, 23
    sc = Scenario(N, seed=seed, full_bodies=True)
    OracleMDP.body_rows = sc.body_indices
    OracleMDP.instances.clear()
    monkeypatch.setattr(terms, "AllstepsMDP", OracleMDP)
    st0 = sc.initial_mdp_state()
    direct = ao.AllstepsOracle(sc.cfg, N, sc.env_origins, sc.joint_limits, sc.body_indices, sc.stone_uniforms(0))
    install_mdp_state(direct, st0)
    phys = sc.physics(direct.steps_pos, direct.curr_target_index, direct.swing_leg)
    scene, robot, left, right = _world(sc, phys, M)

    class Env:  # what `step` / `_reset_idx` touch on `self`
        step = methods["step"]
        _reset_idx = methods["_reset_idx"]

    env = Env()
    env.num_envs, env.device, env.scene = N, "cpu", scene
    env.common_step_counter, env._sim_step_counter = 0, 0
    env.step_dt, env.physics_dt = sc.cfg.step_dt, sc.cfg.step_dt / 4
    env.max_episode_length_s = sc.cfg.episode_length_s
    env.episode_length_buf = st0["episode_length_buf"].clone()
    env.extras = {}
    env.cfg = types.SimpleNamespace(decimation=4, rerender_on_reset=False, sim=types.SimpleNamespace(render_interval=4))
    env.sim = _Quiet()
    env.recorder_manager = _Quiet()
    env.command_manager = _Quiet()

    class Actions(_Quiet):  # ActionManager: `process_action` keeps the raw actions the terms read (`.action`)
        action = torch.zeros(N, 21)

        def process_action(self, a):
            self.action = a

    env.action_manager = Actions()
    pending = {}

    def scene_update(dt=None):  # the last sub-step of the physics loop brings the step's synthetic state
        env._substeps = getattr(env, "_substeps", 0) + 1
        if env._substeps % 4 == 0:
            world = {k: v.clone() for k, v in pending["phys"].items()}
            robot.load_physics(world)
            left.data.force_matrix_w = world["force_matrix_left"]
            right.data.force_matrix_w = world["force_matrix_right"]

    scene.update = scene_update
    scene.write_data_to_sim = lambda: None

    cfgs = manager_cfg.build_manager_cfgs(M)
    terms.binding(env, seed=seed)
    mdp = OracleMDP.instances[0]
    env.observation_manager = M.ObservationManager(cfgs["observations"], env)
    env.reward_manager = M.RewardManager(cfgs["rewards"], env)
    env.termination_manager = M.TerminationManager(cfgs["terminations"], env)
    env.event_manager = M.EventManager(cfgs["events"], env)
    env.curriculum_manager = M.CurriculumManager(cfgs["curriculum"], env)
    assert mdp.counter == 0 and not mdp.pass1_done, "building the managers must not step the MDP"

    # the stand-in's oracle starts from the same mid-episode state as the directly stepped one; its Philox step counter
    # is what the library's would be after the initial reset of a real run
    mdp.orc.load_physics(mdp._load(terms.binding(env).views(env)))
    install_mdp_state(mdp.orc, st0)
    mdp.counter = 1

    n_reset = 0
    for step in range(8):
        phys = sc.physics(direct.steps_pos, direct.curr_target_index, direct.swing_leg)
        m, n = sc.reset_uniforms(step + 2)
        o_obs, o_rew, o_term, o_to, o_ids = direct.step(phys, phys["actions"], m, n, None)
        pending["phys"] = phys
        obs, rew, terminated, time_outs, extras = env.step(phys["actions"].clone())
        assert env._sim_step_counter == 4 * (step + 1) and env.common_step_counter == step + 1
        assert torch.equal(terminated, o_term) and torch.equal(time_outs, o_to), f"step {step}"
        assert torch.allclose(rew, o_rew, rtol=1e-6, atol=1e-6), f"step {step}"
        assert torch.equal(obs["policy"], o_obs), f"step {step}"
        assert torch.equal(env.episode_length_buf, direct.episode_length_buf), f"step {step}"
        if len(o_ids):
            n_reset += len(o_ids)
            assert torch.equal(robot.rec.calls["joint_state"][2], o_ids)
            assert torch.equal(robot.rec.calls["joint_state"][0], direct.reset_writes["joint_pos"])
            assert "Curriculum/allsteps_level/level" in extras["log"]
    assert n_reset > 0


def test_the_reference_direct_rl_env_step_drives_the_hooks(monkeypatch):
    """B1 the same way: the bodies of the reference's own `DirectRLEnv.reset`, `.step` and `._reset_idx`
    (direct_rl_env.py:256-294,296-383,563-584) run on an env whose six hooks are `AllstepsHooksB200` -- the order of
    `_pre_physics_step` / 4 x `_apply_action` / `_get_dones` / `_get_rewards` / `_reset_idx` / `_get_observations`, the
    `.nonzero()`, the counters and `super()._reset_idx` are the reference's, not a restatement.  (CPU: the handle behind
    the hooks is the stand-in over the oracle; the hooks on the real handle are compared in tests/test_gpu_faces.py.)"""
    from allsteps_isaaclab_b200 import env as env_mod
    from oracle import allsteps_oracle as ao
    from oracle import ref_managers
    from scenario import Scenario, install_mdp_state

    methods = ref_managers.load_direct_env_methods()
    N, seed = 72, 31
    sc = Scenario(N, seed=seed, full_bodies=True)
    OracleMDP.body_rows = sc.body_indices
    OracleMDP.instances.clear()
    monkeypatch.setattr(env_mod, "AllstepsMDP", OracleMDP)
    st0 = sc.initial_mdp_state()
    direct = ao.AllstepsOracle(sc.cfg, N, sc.env_origins, sc.joint_limits, sc.body_indices, sc.stone_uniforms(0))
    phys = sc.physics(direct.steps_pos, direct.curr_target_index, direct.swing_leg)
    scene, robot, left, right = _world(sc, phys, None)

    class RefDirectRLEnv:  # the reference's driver, nothing else
        reset, step, _reset_idx = methods["reset"], methods["step"], methods["_reset_idx"]

    class Env(env_mod.AllstepsHooksB200, RefDirectRLEnv):
        pass

    env = Env()
    env.robot, env.sensor_left, env.sensor_right, env.scene = robot, left, right, scene
    env.num_envs, env.device = N, "cpu"
    env.common_step_counter, env._sim_step_counter = 0, 0
    env.step_dt, env.physics_dt = sc.cfg.step_dt, sc.cfg.step_dt / 4
    env.episode_length_buf = torch.zeros(N, dtype=torch.long)  # DRL:179-182
    env.reset_terminated = torch.zeros(N, dtype=torch.bool)
    env.reset_time_outs = torch.zeros(N, dtype=torch.bool)
    env.reset_buf = torch.zeros(N, dtype=torch.bool)
    env.extras = {}
    env.cfg = types.SimpleNamespace(decimation=4, rerender_on_reset=False, wait_for_textures=False, events=None,
                                    action_noise_model=None, observation_noise_model=None,
                                    sim=types.SimpleNamespace(render_interval=4))
    env.sim = _Quiet()
    pending = {}
    calls = []

    def scene_update(dt=None):
        env._substeps = getattr(env, "_substeps", 0) + 1
        if env._substeps % 4 == 0 and "phys" in pending:
            world = {k: v.clone() for k, v in pending["phys"].items()}
            robot.load_physics(world)
            left.data.force_matrix_w = world["force_matrix_left"]
            right.data.force_matrix_w = world["force_matrix_right"]

    scene.update = scene_update
    scene.write_data_to_sim = lambda: calls.append("write_data_to_sim")
    env._init_allsteps_b200(sc.cfg, seed)
    mdp = OracleMDP.instances[0]
    assert torch.equal(mdp.orc.steps_pos, direct.steps_pos), "stones generated by the hook (ENV:71)"

    # ---- DirectRLEnv.reset(): `_reset_idx(all ids)` before any step, then the observations
    fresh = ao.AllstepsOracle(sc.cfg, N, sc.env_origins, sc.joint_limits, sc.body_indices, sc.stone_uniforms(0))
    m, n = sc.reset_uniforms(1)
    fresh.load_physics(phys)
    fresh.reset_rows(torch.arange(N), m, n)
    mdp.orc.load_physics(mdp._load(env._physics_views()))
    obs, extras = env.reset()
    assert torch.equal(obs["policy"], fresh.observations()), "observations of the initial reset"
    assert torch.equal(robot.rec.calls["joint_state"][0], fresh.reset_writes["joint_pos"])
    assert int(env.episode_length_buf.abs().sum()) == 0 and mdp.counter == 1

    # ---- steps from a mid-episode state
    install_mdp_state(direct, st0)
    install_mdp_state(mdp.orc, st0)
    env.episode_length_buf[:] = st0["episode_length_buf"]
    n_reset = 0
    for step in range(8):
        phys = sc.physics(direct.steps_pos, direct.curr_target_index, direct.swing_leg)
        m, n = sc.reset_uniforms(step + 2)
        direct.clamp_actions(phys["actions"])
        o_eff = direct.joint_efforts().clone()  # ENV:270-274 at the levels the step starts with
        o_obs, o_rew, o_term, o_to, o_ids = direct.step(phys, phys["actions"], m, n, None)
        pending["phys"] = phys
        before = mdp.launch_count
        obs, rew, terminated, time_outs, extras = env.step(phys["actions"].clone())
        assert env._sim_step_counter == 4 * (step + 1) and env.common_step_counter == step + 1
        assert torch.equal(terminated, o_term) and torch.equal(time_outs, o_to), f"step {step}"
        assert torch.equal(env.reset_buf, o_term | o_to)
        assert torch.allclose(rew, o_rew, rtol=1e-6, atol=1e-6), f"step {step}"
        assert torch.equal(obs["policy"], o_obs), f"step {step}"
        assert torch.equal(env.episode_length_buf, direct.episode_length_buf), f"step {step}"
        # four `_apply_action` calls of the decimation loop: one launch, the efforts of the levels BEFORE this step's reset
        assert torch.allclose(robot.rec.calls["joint_effort_target"][0], o_eff, rtol=1e-6, atol=1e-6), f"step {step}"
        # launches: efforts 1 + pass 1 + (reset + pass 2 | nothing)
        assert mdp.launch_count - before == (4 if len(o_ids) else 2), f"step {step}"
        if len(o_ids):
            n_reset += len(o_ids)
            assert torch.equal(robot.rec.calls["joint_state"][2], o_ids)
            assert torch.equal(robot.rec.calls["joint_state"][0], direct.reset_writes["joint_pos"])
    assert n_reset > 0
