"""The rollout loop around the engine: ksim ``RLTask._single_unroll`` (``jax.lax.scan`` over the control steps of a chunk with
``done -> engine.reset`` and one ``Trajectory`` entry per step; ksim 0.1.99 ksim/task/rl.py [R], un-vendored:
/root/reference/uv.lock:1872-1873; its user in the reference: /root/reference/train.py:1727, 1737-1763) — SURVEY.md §8a row a19.

Here every control step is ONE launch (zb_step_obs_f32: S fused substeps + terminations + partial reset + observation epilogue)
whose output pointers are slices of the time-major trajectory arrays, so the launch itself is the trajectory writer: no gather,
no copy.  With the actions given up front (random-action rollouts, BASELINE.json configs[1-2]) the T launches of a chunk are
captured into one CUDA graph.  With the reference's actor in the loop (``unroll_actor``: ``ActorPolicy`` = zb_actor_f32, SURVEY.md §8f
rank 4) a control step is TWO launches — actor (reads ``trajectory.actor_obs[t]``, writes ``trajectory.action[t]``, carries its GRU state,
restarts it where ``done[t-1]``) and physics — and the 2T launches of a chunk are again one CUDA graph: the whole ``jax.lax.scan`` of
``RLTask._single_unroll`` [R] runs on the device without the host.  Any other ``policy`` callable is issued step by step.

Index alignment (as ksim's rollout stores it [R]):
  state-like  [T, N, ...]    qpos, qvel, xpos, xquat, done : `next_physics_state` of step t BEFORE the done -> reset select (the
                             terminal state where an episode ends; ZbAutoResetArgs.term_*)
  obs-like    [T+1, N, ...]  actor_obs, critic_obs, feet_pos_obs, reward_terms, sensordata : slot t = what the action of step t
                             was computed from (written by the launch of step t-1, slot 0 by the reset launch / the previous
                             chunk); slot T = first observation of the next chunk
  action      [T, N, 20]
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass

import torch

from . import model as M
from . import rewards as R
from .engine import NUM_ACTOR_INPUTS, NUM_CRITIC_INPUTS, NUM_REWARD_TERMS, EnvParams, Observations, PhysicsState, ZbotCudaEngine


@dataclass
class Trajectory:
    """Duck-types ``ksim.Trajectory`` for the fields the reference reads (train.py:297, 476, 681: qpos / qvel / xpos / xquat;
    obs[...]; action; done), time-major."""

    qpos: torch.Tensor
    qvel: torch.Tensor
    xpos: torch.Tensor
    xquat: torch.Tensor
    done: torch.Tensor
    action: torch.Tensor
    actor_obs: torch.Tensor
    critic_obs: torch.Tensor
    feet_pos_obs: torch.Tensor
    reward_terms: torch.Tensor
    sensordata: torch.Tensor

    @staticmethod
    def allocate(T: int, n: int, device) -> "Trajectory":
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)  # noqa: E731
        return Trajectory(qpos=z(T, n, M.NQ), qvel=z(T, n, M.NV), xpos=z(T, n, M.NBODY, 3), xquat=z(T, n, M.NBODY, 4), done=z(T, n, dt=torch.uint8),
                          action=z(T, n, M.NU), actor_obs=z(T + 1, n, NUM_ACTOR_INPUTS), critic_obs=z(T + 1, n, NUM_CRITIC_INPUTS),
                          feet_pos_obs=z(T + 1, n, 6), reward_terms=z(T + 1, n, NUM_REWARD_TERMS), sensordata=z(T + 1, n, M.NSENSORDATA))

    OBS_LIKE = ("actor_obs", "critic_obs", "feet_pos_obs", "reward_terms", "sensordata")

    def bytes_per_env_step(self) -> int:
        """Trajectory bytes written per environment per control step."""
        return 4 * (M.NQ + M.NV + M.NBODY * 7 + NUM_ACTOR_INPUTS + NUM_CRITIC_INPUTS + 6 + NUM_REWARD_TERMS + M.NSENSORDATA) + 1


class Rollout:
    """T-step rollout chunks of an engine's whole batch, written straight into a ``Trajectory``."""

    def __init__(self, eng: ZbotCudaEngine, T: int, physics_model: EnvParams | None = None, n: int | None = None):
        self.eng, self.T, self.ep = eng, int(T), physics_model
        self.n = eng.n_envs if n is None else n
        dev = eng.device
        self.traj = Trajectory.allocate(self.T, self.n, dev)
        self.rq, self.rv, self.rt = eng.sample_reset(self.n, physics_model)
        self.reset_imu_lag = eng.sample_imu_lag(self.n, (0.0, 0.1))
        self.carry_obs = eng.new_observations(self.n, critic=False)  # owns the filter carry (imu_x, imu_lag); its actor_obs is unused
        self.airtime_carry = torch.zeros((self.n, 2), dtype=torch.float32, device=dev)
        self.d_push = torch.zeros((self.n, 7), dtype=torch.float32, device=dev) if eng.push is not None else None
        eng.outputs = tuple(k for k in eng.outputs if k in ("sensordata", "contact_mask", "solver_niter", "flags"))  # the rest lives in the trajectory
        self.ps = eng.new_state(self.n)
        self._graph = None
        self._plans = None
        self._actor_graphs = {}
        self.prev_done = torch.zeros((self.n,), dtype=torch.uint8, device=dev)  # `done` of the last step of the previous chunk
        self.noise_u = self.noise_eps = self.pushes = None

    def _obs_slot(self, t: int) -> Observations:
        tr, c = self.traj, self.carry_obs
        return Observations(imu_x=c.imu_x, imu_lag=c.imu_lag, actor_obs=tr.actor_obs[t], critic_obs=tr.critic_obs[t], lag_range=c.lag_range,
                            feet_pos_obs=tr.feet_pos_obs[t], reward_terms=tr.reward_terms[t])

    def _state_slot(self, t: int) -> PhysicsState:
        """The engine state with the obs-like mjx.Data output (sensordata) redirected to slot t of the trajectory."""
        return dataclasses.replace(self.ps, data=dataclasses.replace(self.ps.data, sensordata=self.traj.sensordata[t]))

    def reset(self) -> None:
        """MjxEngine.reset for the whole batch; the first observation goes to slot 0."""
        self.eng.reset(self.ep, physics_state=self._state_slot(0), reset_qpos=self.rq, reset_qvel=self.rv, reset_t_rem=self.rt,
                       obs=self._obs_slot(0), reset_imu_lag=self.reset_imu_lag)
        self.airtime_carry.zero_()

    def plans(self):
        """One pre-bound control step per time index (pointers into slice t / t+1 of the trajectory)."""
        if self._plans is None:
            tr = self.traj
            self._plans = [self.eng.plan_step(self._state_slot(t + 1), self.ep, auto_reset=(self.rq, self.rv, self.rt), done_out=tr.done[t],
                                              obs=self._obs_slot(t + 1), reset_imu_lag=self.reset_imu_lag, track_action=False,
                                              terminal=(tr.qpos[t], tr.qvel[t], tr.xpos[t], tr.xquat[t])) for t in range(self.T)]
        return self._plans

    def unroll(self, policy=None, pushes: torch.Tensor | None = None) -> Trajectory:
        """One chunk.  ``policy(actor_obs [N, 50]) -> action [N, 20]`` is called per step; with ``policy=None`` the actions already in
        ``trajectory.action`` are used and the whole chunk replays as ONE CUDA graph.  ``pushes`` [T, N, 7]: pre-sampled PushEvent
        numbers (``eng.sample_push``)."""
        tr, plans = self.traj, self.plans()
        if policy is None and pushes is None and self.d_push is None:
            if self._graph is None:
                s = torch.cuda.Stream(device=self.eng.device)
                s.wait_stream(torch.cuda.current_stream(self.eng.device))
                with torch.cuda.stream(s):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=s):
                        for t in range(self.T):
                            plans[t](tr.action[t], None)
                torch.cuda.current_stream(self.eng.device).wait_stream(s)
                self._graph = g
            else:
                self._graph.replay()
            return tr
        for t in range(self.T):
            if policy is not None:
                tr.action[t].copy_(policy(tr.actor_obs[t]))
            if self.d_push is not None:
                self.d_push.copy_(pushes[t] if pushes is not None else self.eng.sample_push(self.n))
            plans[t](tr.action[t], self.d_push)
        return tr

    def unroll_actor(self, policy, argmax: bool = True, graph: bool = True) -> Trajectory:
        """One chunk with the reference's actor in the loop (``sample_action``, train.py:1737-1763): per control step
        ``policy.enqueue`` (zb_actor_f32) then the physics launch, nothing else — no host round trip, no copy.  ``argmax=False``
        samples with the numbers in ``self.noise_u`` / ``self.noise_eps`` [T, N, 20] (refill them in place before every chunk:
        ``sample_noise()``); a configured PushEvent reads ``self.pushes`` [T, N, 7] (``sample_pushes()``).  The carry of an
        environment restarts at the step after its ``done`` (``prev_done`` carries that across chunks).  With ``graph=True`` the 2T
        launches are captured once per (argmax) and replayed."""
        tr, plans, T, dev = self.traj, self.plans(), self.T, self.eng.device
        if not argmax and self.noise_u is None:
            self.noise_u = torch.zeros((T, self.n, M.NU), dtype=torch.float32, device=dev)
            self.noise_eps = torch.zeros((T, self.n, M.NU), dtype=torch.float32, device=dev)
            self.sample_noise()
        if self.eng.push is not None and self.pushes is None:
            self.pushes = torch.zeros((T, self.n, 7), dtype=torch.float32, device=dev)
            self.sample_pushes()

        def issue():
            for t in range(T):
                policy.enqueue(tr.actor_obs[t], tr.action[t], carry_reset=self.prev_done if t == 0 else tr.done[t - 1],
                               noise_u=None if argmax else self.noise_u[t], noise_eps=None if argmax else self.noise_eps[t])
                plans[t](tr.action[t], None if self.pushes is None else self.pushes[t])
            self.prev_done.copy_(tr.done[T - 1])

        if not graph:
            issue()
            return tr
        key = (id(policy), bool(argmax))
        if key not in self._actor_graphs:
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    issue()
            torch.cuda.current_stream(dev).wait_stream(s)
            self._actor_graphs[key] = g
        self._actor_graphs[key].replay()
        return tr

    def sample_noise(self, generator: torch.Generator | None = None) -> None:
        """Refill the action-sampling numbers of a chunk in place: U[0,1) for the mixture component, N(0,1) for the Gaussian."""
        self.noise_u.uniform_(0.0, 1.0, generator=generator)
        self.noise_eps.normal_(0.0, 1.0, generator=generator)

    def sample_pushes(self) -> None:
        for t in range(self.T):
            self.pushes[t].copy_(self.eng.sample_push(self.n))

    def next_chunk(self) -> None:
        """Carry the last observation slot over to slot 0 (five small device copies per chunk)."""
        for k in Trajectory.OBS_LIKE:
            a = getattr(self.traj, k)
            a[0].copy_(a[self.T])

    def rewards(self) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
        """The reference's reward list over the chunk (train.py:1546-1586): ([T, N] total, {name: [T, N]} raw terms).
        FeetAirtimeReward runs as one kernel over the chunk's contact flags (zb_feet_airtime_f32), its carry lives here."""
        tr, T = self.traj, self.T
        air = self.eng.feet_airtime_reward(tr.reward_terms[:T, :, 4], tr.reward_terms[:T, :, 5], tr.done, self.airtime_carry, touchdown_penalty=0.3)
        m = self.eng.model
        terms = R.reward_terms(tr, air, int(m.site_body[1]), int(m.site_body[2]))
        return R.total(terms), terms
