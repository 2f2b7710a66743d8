"""PPO on the batched engine (BASELINE.json configs[2]: N envs x T-step rollout with GAE on GPU).

Shape of the reference trainer, ppo/train.py:224-335, re-done in PyTorch: the policy network and
its optimiser are ordinary torch modules (they are not on the accelerated path and are not held to
parity); everything between two policy forward passes runs in the CUDA library:

    get_rollouts (ppo/train.py:156-197)  -> env.step_discrete: action decode + step + observation
    roll.record / roll.batched           -> DeviceRollout: time-major device buffers
    rl.make_returns / rl.gae (:268-277)  -> one agar_gae launch over [T, N*A]
    observations kept for the SGD phase  -> state snapshots (layout.words floats per arena-step
                                            instead of A*C*G*G) re-rasterised per minibatch with
                                            agar_observe_records (SURVEY.md §7 "hard parts")

Multi-GPU: one process per GPU, each with its own arena shard (sharding.py); the only collective
is the gradient all-reduce of DistributedDataParallel over NCCL.
"""
import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .env import BatchedAgarioEnv
from .gae import gae_and_returns
from .rollout import DeviceRollout


class ActorCritic(nn.Module):
    """ppo/train.py:33-76: Conv(16,5x5) SiLU Conv(16,5x5) SiLU Conv(5,3x3) SiLU Dropout Conv(5,3x3)
    SiLU Dropout Flatten Dense(num_features), then Dense(num_actions) and Dense(1) heads.  stax's
    `Dropout(0.95)` is a keep rate, i.e. p = 0.05 here; stax.Conv pads VALID."""

    def __init__(self, obs_shape, num_actions: int, num_features: int = 1024, p_drop: float = 0.05):
        super().__init__()
        c, g, _ = obs_shape
        self.features = nn.Sequential(
            nn.Conv2d(c, 16, 5), nn.SiLU(), nn.Conv2d(16, 16, 5), nn.SiLU(),
            nn.Conv2d(16, 5, 3), nn.SiLU(), nn.Dropout(p_drop), nn.Conv2d(5, 5, 3), nn.SiLU(), nn.Dropout(p_drop),
            nn.Flatten())
        side = g - 4 - 4 - 2 - 2
        self.dense = nn.Linear(5 * side * side, num_features)
        self.actor = nn.Linear(num_features, num_actions)
        self.critic = nn.Linear(num_features, 1)

    def forward(self, obs):
        h = self.dense(self.features(obs))
        return self.actor(h), self.critic(h)[:, 0]


def ppo_actor_loss(logits, actions, advantages, old_logits, clip_eps):
    """rl/__init__.py:12-46."""
    new_lp = F.log_softmax(logits, 1).gather(1, actions[:, None])[:, 0]
    old_lp = F.log_softmax(old_logits, 1).gather(1, actions[:, None])[:, 0]
    ratio = torch.exp(new_lp - old_lp)
    adv = advantages.detach()
    return -torch.minimum(adv * ratio, adv * ratio.clamp(1 - clip_eps, 1 + clip_eps)).mean()


def critic_loss(values, returns, delta: float = 1.0):
    """rl/__init__.py:49-57 (Huber)."""
    a = values - returns
    return torch.where(a.abs() < delta, a.square() / 2, delta * (a.abs() - delta / 2)).mean()


def policy_entropy(logits):
    """rl.xs(softmax(logits), logits).mean() of ppo/train.py:125."""
    lp = F.log_softmax(logits, 1)
    return -(lp.exp() * lp).sum(1).mean()


@dataclass
class PPOConfig:
    episode_length: int = 128        # T (configs[2]); the reference uses 512 (ppo.textproto:31)
    num_sgd_steps: int = 8           # ppo.textproto:33
    batch_size: int = 128            # ppo.textproto:34
    gamma: float = 0.75              # ppo.textproto:62
    lam: float = 0.1                 # ppo.textproto:67
    clip_eps: float = 0.2            # ppo.textproto:71
    num_features: int = 1024         # ppo.textproto:26
    lr_base: float = 1e-3            # exponential_decay base (ppo.textproto:38-58)
    lr_decay_steps: float = 500.0    # base * e^(-t/500) (ppo/train.py:210-213)
    lr_shift: tuple = (1e-5, 1e-5, 1e-4)   # feature extractor, actor, critic
    lr_scale: tuple = (1.0, 1.0, 10.0)
    grad_clip: float = 1.0           # optimizers.clip_grads(dparams, 1), ppo/train.py:331
    # What the SGD phase reads its observations from.  "presampled" (default): the minibatch rows of the coming SGD
    # phase are drawn before the rollout (they are uniform over the rollout and do not depend on its content) and only
    # those observations are kept - no snapshot traffic at all.  "state": a state snapshot per step, re-rasterised per
    # minibatch with agar_observe_records (any row can be read afterwards; 1.26 GB per step at 65,536 arenas).
    # "observations": the [T, B, C, G, G] observations themselves (412 GB at configs[2] size).
    keep: str = "presampled"
    presample_factor: int = 4        # candidates drawn per minibatch row (rows of finished agents are rejected)
    keep_observations: bool = False  # deprecated spelling of keep="observations"
    amp_bf16: bool = False           # policy network under torch.autocast(bfloat16) + channels_last (the policy is
                                     # plain PyTorch and outside the parity scope; the env path stays float32)


class PPOTrainer:
    """ppo/train.py:224-335 on the batched engine.

    Two episode regimes, chosen by the env's `auto_reset`:
      auto_reset = 0 (the reference's, ppo/train.py:156-197, 268-290): every episode starts with env.reset(); the
        rollout records the PRE-step dones (:181), stops early once every agent is done (:166), the bootstrap values of
        finished agents are zeroed (:189-195), returns and GAE run over whole columns without a cut (:272-277) and the
        rows of finished agents are dropped from the SGD batch (:284-290).
      auto_reset = 1 (vector-env regime for configs[2]): arenas restart inside the step in which their agents finish,
        episodes continue across train_episode() calls, the POST-step dones cut the GAE recursion.
    ddp=True wraps the policy in DistributedDataParallel: one process per GPU, each with its own arena shard
    (sharding.py); the gradient all-reduce over NCCL is the only collective of the whole path."""

    def __init__(self, env: BatchedAgarioEnv, cfg: PPOConfig = PPOConfig(), seed: int = 0, ddp: bool = False):
        self.env, self.cfg = env, cfg
        self.dev = env.device
        self.keep = "observations" if cfg.keep_observations else cfg.keep
        if self.keep not in ("presampled", "state", "observations"):
            raise ValueError("PPOConfig.keep must be 'presampled', 'state' or 'observations'")
        self.reference_episodes = not env.cfg.auto_reset
        torch.manual_seed(seed)   # identical initial parameters on every rank
        self.model = ActorCritic(env.observation_space.shape, env.num_actions, cfg.num_features).to(self.dev)
        if cfg.amp_bf16:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.ddp = bool(ddp)
        self.net = nn.parallel.DistributedDataParallel(self.model, device_ids=[self.dev.index]) if ddp else self.model
        groups = [list(self.model.features.parameters()) + list(self.model.dense.parameters()),
                  list(self.model.actor.parameters()), list(self.model.critic.parameters())]
        self.opt = torch.optim.Adam([{"params": g, "lr": 1.0} for g in groups])
        self.sched = torch.optim.lr_scheduler.LambdaLR(
            self.opt, [lambda t, i=i: cfg.lr_scale[i] * cfg.lr_base * math.exp(-t / cfg.lr_decay_steps) + cfg.lr_shift[i]
                       for i in range(3)])
        rank = torch.distributed.get_rank() if (ddp and torch.distributed.is_initialized()) else 0
        self.gen = torch.Generator(device=self.dev); self.gen.manual_seed(seed + 7919 * rank)   # sampling differs per shard
        self.B = env.num_envs * env.num_agents
        self.roll = DeviceRollout(cfg.episode_length, device=self.dev)
        self.obs = env.reset()
        self.sgd_steps = 0
        self.episodes = 0
        self._mb = None

    def _flat(self, obs):
        return obs.reshape((self.B,) + tuple(obs.shape[2:]))

    def _policy(self, obs):
        """logits, values in float32 (optionally evaluated under bf16 autocast)"""
        if not self.cfg.amp_bf16:
            return self.net(obs)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, values = self.net(obs.contiguous(memory_format=torch.channels_last))
        return logits.float(), values.float()

    # ---- minibatch rows drawn ahead of the rollout (keep = "presampled") ----------------------
    def _presample(self):
        """For each SGD step, `presample_factor * batch_size` candidate rows (t, b), uniform over the rollout; their
        observations are gathered while the rollout runs.  The SGD step takes the first batch_size candidates whose row
        counts (all of them in the auto-reset regime; in the reference regime rows of finished agents are rejected,
        which leaves the draw uniform over the kept rows, ppo/train.py:284-290 and :316-317)."""
        cfg, T = self.cfg, self.cfg.episode_length
        n = cfg.num_sgd_steps * cfg.batch_size * (cfg.presample_factor if self.reference_episodes else 1)
        n = min(n, T * self.B)
        pick = torch.randperm(T * self.B, generator=self.gen, device=self.dev)[:n]
        t_idx, b_idx = pick // self.B, pick % self.B
        order = torch.argsort(t_idx, stable=True)
        counts = torch.bincount(t_idx, minlength=T).cpu()
        offs = torch.zeros(T + 1, dtype=torch.long); offs[1:] = counts.cumsum(0)
        obs_shape = tuple(self.env.obs_shape[2:])
        self._mb = {"t": t_idx, "b": b_idx, "order": order, "offs": offs.tolist(),
                    "obs": torch.empty((n,) + obs_shape, dtype=torch.float32, device=self.dev)}

    def _presample_gather(self, t, flat_obs):
        mb = self._mb
        lo, hi = mb["offs"][t], mb["offs"][t + 1]
        if hi > lo:
            slots = mb["order"][lo:hi]
            mb["obs"][slots] = flat_obs[mb["b"][slots]]

    @torch.no_grad()
    def collect(self):
        """get_rollouts (ppo/train.py:156-197) for all arenas at once."""
        env, T = self.env, self.cfg.episode_length
        self.roll.clear()
        self.model.train()  # the reference always builds the extractor with mode='train' (App. D)
        if self.reference_episodes:
            self.obs = env.reset()                                   # ppo/train.py:164
        dones = torch.zeros(self.B, dtype=torch.uint8, device=self.dev)
        if self.keep == "presampled":
            self._presample()
        check_every = 1 if self.B <= 4096 else 32                    # the all-done test costs a host sync
        for t in range(T):
            flat = self._flat(self.obs)
            logits, values = self._policy(flat)
            actions = torch.multinomial(F.softmax(logits, 1), 1, generator=self.gen)[:, 0]
            rec = {"actions": actions, "values": values, "action_logits": logits}
            if self.keep == "observations":
                rec["observations"] = flat
            elif self.keep == "state":
                rec["state"] = env.get_state()
            else:
                self._presample_gather(t, flat)
            self.obs, reward, done, _ = env.step_discrete(actions.view(env.num_envs, env.num_agents).to(torch.int32))
            rec["rewards"] = reward.reshape(self.B)
            # the reference records the dones known BEFORE the step (ppo/train.py:181); the auto-reset regime needs the
            # post-step ones (they mark where an arena restarted)
            rec["dones"] = dones if self.reference_episodes else done.reshape(self.B)
            self.roll.record(rec)
            dones = done.reshape(self.B)
            if self.reference_episodes and (t % check_every == check_every - 1) and bool(dones.all()):
                break                                                # ppo/train.py:166
        _, boot = self._policy(self._flat(self.obs))
        if self.reference_episodes:
            boot = boot.masked_fill(dones.bool(), 0.0)               # ppo/train.py:189-195
        self.roll.record({"values": boot})  # bootstrap row
        self.episodes += 1

    def advantages(self):
        """rl.make_returns + rl.gae for every column in one launch (ppo/train.py:268-277).  Reference regime: no cut
        inside the columns (the reference has none; finished agents' bootstrap is zero and their rows are dropped later).
        Auto-reset regime: done[t] cuts the recursion where an arena restarted."""
        r, v = self.roll["rewards"], self.roll["values"]
        d = None if self.reference_episodes else self.roll["dones"]
        return gae_and_returns(r, v, self.cfg.gamma, self.cfg.lam, dones=d, end_value=v[-1].contiguous())

    def _observations(self, t_idx, b_idx):
        if self.keep == "observations":
            return self.roll["observations"][t_idx, b_idx]
        A = self.env.num_agents
        rec = self.roll["state"][t_idx, b_idx // A]                    # [mb, words]
        obs = self.env.observe_records(rec)                            # [mb, A, C, G, G]
        return obs[torch.arange(len(b_idx), device=self.dev), b_idx % A]

    def _minibatch(self, k):
        """(t, b, observations) of SGD step k"""
        cfg, T = self.cfg, len(self.roll["rewards"])
        if self.keep == "presampled":
            mb = self._mb
            per = len(mb["t"]) // cfg.num_sgd_steps
            sl = slice(k * per, (k + 1) * per)
            t_idx, b_idx, obs = mb["t"][sl], mb["b"][sl], mb["obs"][sl]
            ok = t_idx < T                                             # an early exit leaves later rows unrecorded
            if self.reference_episodes:
                ok &= self.roll["dones"][t_idx.clamp(max=T - 1), b_idx] == 0
            keep = ok.nonzero()[:, 0][:cfg.batch_size]
            return t_idx[keep], b_idx[keep], obs[keep]
        if self.reference_episodes:                                    # rows with ~done, ppo/train.py:284-290
            valid = (self.roll["dones"].reshape(-1) == 0).nonzero()[:, 0]
            pick = valid[torch.randperm(len(valid), generator=self.gen, device=self.dev)[:cfg.batch_size]]
        else:
            pick = torch.randperm(T * self.B, generator=self.gen, device=self.dev)[:cfg.batch_size]
        t_idx, b_idx = pick // self.B, pick % self.B
        return t_idx, b_idx, self._observations(t_idx, b_idx)

    def update(self, adv, ret):
        cfg = self.cfg
        stats = {}
        for k in range(cfg.num_sgd_steps):
            t_idx, b_idx, obs = self._minibatch(k)
            if len(t_idx) == 0:
                continue
            logits, values = self._policy(obs)
            la = ppo_actor_loss(logits, self.roll["actions"][t_idx, b_idx], adv[t_idx, b_idx],
                                self.roll["action_logits"][t_idx, b_idx], cfg.clip_eps)
            lc = critic_loss(values, ret[t_idx, b_idx])
            ent = policy_entropy(logits)
            loss = la + lc + 1.0 / (ent * 100.0)   # the reference's entropy term, ppo/train.py:101
            self.opt.zero_grad(set_to_none=True)
            loss.backward()                        # DDP all-reduces the gradients over NCCL here
            nn.utils.clip_grad_norm_(self.model.parameters(), cfg.grad_clip)
            self.opt.step(); self.sched.step()
            self.sgd_steps += 1
            stats = {"loss": loss.item(), "actor": la.item(), "critic": lc.item(), "entropy": ent.item()}
        return stats

    def train_episode(self):
        self.collect()
        adv, ret = self.advantages()
        stats = self.update(adv, ret)
        stats["mean_return"] = float(self.roll["rewards"].sum(0).mean())
        return stats

    # ---- checkpoint / resume (the reference saves its parameters per episode, a2c/training.py:151-153) ----
    def state_dict(self):
        """Everything the next episode depends on: parameters, optimiser and schedule, the engine's world state and
        seed (its RNG is counter-based: the tick counter lives in the state record), the sampling generator and the
        global torch RNG states (dropout)."""
        return {
            "model": self.model.state_dict(), "opt": self.opt.state_dict(), "sched": self.sched.state_dict(),
            "env_state": self.env.get_state().cpu(), "env_seed": int(self.env.seed_value), "obs": self.obs.cpu(),
            "gen": self.gen.get_state(), "torch_rng": torch.get_rng_state(), "cuda_rng": torch.cuda.get_rng_state(self.dev),
            "sgd_steps": self.sgd_steps, "episodes": self.episodes, "cfg": dict(self.cfg.__dict__),
        }

    def load_state_dict(self, sd):
        self.model.load_state_dict(sd["model"])
        self.opt.load_state_dict(sd["opt"])
        self.sched.load_state_dict(sd["sched"])
        self.env.seed(sd["env_seed"])          # (restarts the tick counters; the state record below restores them)
        self.env.set_state(sd["env_state"].to(self.dev))
        self.obs = sd["obs"].to(self.dev)
        self.gen.set_state(sd["gen"])
        torch.set_rng_state(sd["torch_rng"])
        torch.cuda.set_rng_state(sd["cuda_rng"], self.dev)
        self.sgd_steps, self.episodes = int(sd["sgd_steps"]), int(sd["episodes"])

    def save(self, path):
        torch.save(self.state_dict(), path)

    def resume(self, path):
        self.load_state_dict(torch.load(path, map_location="cpu", weights_only=False))
        return self
