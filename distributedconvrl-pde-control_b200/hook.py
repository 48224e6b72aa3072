"""Host-side mirror of src/PDEhook.jl with batched semantics.

Same fields and stage callbacks as the reference's `PDEhook` (PDEhook.jl:8-103).  With B environments per context:

  * `reward` accumulates the mean over ALL columns of the batch per step (`hook.reward += mean(reward(env))`,
    PDEhook.jl:52) and `rewards_per_env` additionally keeps the per-environment episode returns;
  * `bestNNA` / `currentNNA` are host copies of the behavior actor pulled from the device
    (`copyto!(hook.bestNNA, agent.policy.behavior_actor)`, PDEhook.jl:69,95);
  * `bestDF` / `currentDF` record (timestep, action, p, y, reward) rows of ONE tracked environment
    (`track_env`, default 0) so that `plot_heat(plot_best = true)`-style consumers keep working without
    copying the whole batch every step (`collect_bestDF`, PDEhook.jl:54-62).
"""
import numpy as np


class PDEhook:
    def __init__(self, *, use_random_init=False, collect_history=False, collect_NNA=True, collect_bestDF=True,
                 min_best_episode=0, error_detection=None, generate_random_init=None, track_env=0):
        self.rewards, self.rewards_compare, self.rewards_per_env = [], [], []
        self.reward, self.ep = 0.0, 1
        self.use_random_init, self.collect_history = use_random_init, collect_history
        self.collect_NNA, self.collect_bestDF = collect_NNA, collect_bestDF
        self.min_best_episode = min_best_episode
        self.bestNNA = self.currentNNA = None
        self.bestDF, self.currentDF, self.history = [], [], []
        self.bestreward, self.bestepisode = -1000000.0, 0
        self.errored_episodes = []
        self.error_detection = error_detection or (lambda y: False)
        self.generate_random_init = generate_random_init
        self.track_env = int(track_env)
        self._ret = None

    # PreExperimentStage, PDEhook.jl:35-40
    def pre_experiment(self, env, policy):
        if self.collect_NNA and self.currentNNA is None:
            self.currentNNA = policy.behavior_actor.sync_from_device().copy()
            self.bestNNA = self.currentNNA.copy()

    # PreEpisodeStage, PDEhook.jl:42-49
    def pre_episode(self, env, policy=None):
        if policy is not None:
            self.pre_experiment(env, policy)
        if self.use_random_init and self.generate_random_init is not None:
            env.set_y0(self.generate_random_init(env.n_envs))
            env.reset()
        self._ret = np.zeros(env.n_envs)

    # PostActStage, PDEhook.jl:51-63
    def post_act(self, env):
        r = np.asarray(env.reward, dtype=np.float64)
        self.reward += float(r.mean())
        self._ret += r.reshape(env.n_envs, -1).mean(axis=1)
        if self.collect_bestDF:
            # only the tracked environment's slices cross PCIe (pdeb200_get_env), not the whole batch
            from . import _lib as L
            b = self.track_env
            self.currentDF.append({
                "timestep": int(env.get_env(L.ARR_STEPS, b)[0]),
                "action": env.get_env(L.ARR_ACTION, b).reshape(env.n_actuators, env.a_rows).T.reshape(-1).copy(),
                "p": env.p_of(b).copy(),
                "y": env.y_of(b).copy(),
                "reward": r.reshape(env.n_envs, -1)[b].copy(),
            })

    # PostEpisodeStage, PDEhook.jl:65-97
    def post_episode(self, env, policy):
        finished = bool(np.any(env.time >= env.te))     # the shared clock ran out (diverged envs were reset on the way)
        if finished and self.ep >= self.min_best_episode:
            self.rewards_compare.append(self.reward)
            if self.collect_NNA and self.reward >= max(self.rewards_compare):
                self.bestNNA = policy.behavior_actor.sync_from_device().copy()
                self.bestreward, self.bestepisode = self.reward, self.ep
                if self.collect_bestDF:
                    self.bestDF = list(self.currentDF)
        if not finished:
            if self.error_detection(env.y_of(self.track_env)):
                self.errored_episodes.append(self.ep)
        if self.collect_history:
            self.history.append(self.currentDF)
        self.currentDF = []
        self.ep += 1
        self.rewards.append(self.reward)
        self.rewards_per_env.append(self._ret.copy())
        self.reward = 0.0
        if self.collect_NNA:
            self.currentNNA = policy.behavior_actor.sync_from_device().copy()
