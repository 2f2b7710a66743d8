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
    keep_observations: bool = False  # True: keep [T,B,C,G,G] observations instead of state snapshots
    amp_bf16: bool = False           # policy network under torch.autocast(bfloat16) + channels_last (the policy is
                                     # plain PyTorch and outside the parity scope; the env path stays float32)


class PPOTrainer:
    def __init__(self, env: BatchedAgarioEnv, cfg: PPOConfig = PPOConfig(), seed: int = 0, ddp: bool = False):
        if not env.cfg.auto_reset:
            raise ValueError("PPOTrainer expects auto_reset=1 (arenas restart when all their agents are done)")
        self.env, self.cfg = env, cfg
        self.dev = env.device
        torch.manual_seed(seed)
        self.model = ActorCritic(env.observation_space.shape, env.num_actions, cfg.num_features).to(self.dev)
        if cfg.amp_bf16:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.net = nn.parallel.DistributedDataParallel(self.model, device_ids=[self.dev.index]) if ddp else self.model
        groups = [list(self.model.features.parameters()) + list(self.model.dense.parameters()),
                  list(self.model.actor.parameters()), list(self.model.critic.parameters())]
        self.opt = torch.optim.Adam([{"params": g, "lr": 1.0} for g in groups])
        self.sched = torch.optim.lr_scheduler.LambdaLR(
            self.opt, [lambda t, i=i: cfg.lr_scale[i] * cfg.lr_base * math.exp(-t / cfg.lr_decay_steps) + cfg.lr_shift[i]
                       for i in range(3)])
        self.gen = torch.Generator(device=self.dev); self.gen.manual_seed(seed)
        self.B = env.num_envs * env.num_agents
        self.roll = DeviceRollout(cfg.episode_length, device=self.dev)
        self.obs = env.reset()
        self.sgd_steps = 0

    def _flat(self, obs):
        return obs.reshape((self.B,) + tuple(obs.shape[2:]))

    def _policy(self, obs):
        """logits, values in float32 (optionally evaluated under bf16 autocast)"""
        if not self.cfg.amp_bf16:
            return self.net(obs)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, values = self.net(obs.contiguous(memory_format=torch.channels_last))
        return logits.float(), values.float()

    @torch.no_grad()
    def collect(self):
        """get_rollouts (ppo/train.py:156-197) for all arenas at once."""
        env, T = self.env, self.cfg.episode_length
        self.roll.clear()
        self.model.train()  # the reference always builds the extractor with mode='train' (App. D)
        for t in range(T):
            logits, values = self._policy(self._flat(self.obs))
            actions = torch.multinomial(F.softmax(logits, 1), 1, generator=self.gen)[:, 0]
            rec = {"actions": actions, "values": values, "action_logits": logits}
            rec["observations" if self.cfg.keep_observations else "state"] = (
                self._flat(self.obs) if self.cfg.keep_observations else env.get_state())
            self.obs, reward, done, _ = env.step_discrete(actions.view(env.num_envs, env.num_agents).to(torch.int32))
            rec["rewards"] = reward.reshape(self.B)
            rec["dones"] = done.reshape(self.B)
            self.roll.record(rec)
        _, boot = self._policy(self._flat(self.obs))
        self.roll.record({"values": boot})  # bootstrap row, ppo/train.py:189-195

    def advantages(self):
        """rl.make_returns + rl.gae for every column in one launch (ppo/train.py:268-277); done[t]
        cuts the recursion where an arena restarted (the reference instead zeroes the bootstrap and
        drops done rows, :189-195 and :284-290)."""
        r, v, d = self.roll["rewards"], self.roll["values"], self.roll["dones"]
        return gae_and_returns(r, v, self.cfg.gamma, self.cfg.lam, dones=d, end_value=v[-1].contiguous())

    def _observations(self, t_idx, b_idx):
        if self.cfg.keep_observations:
            return self.roll["observations"][t_idx, b_idx]
        A = self.env.num_agents
        rec = self.roll["state"][t_idx, b_idx // A]                    # [mb, words]
        obs = self.env.observe_records(rec)                            # [mb, A, C, G, G]
        return obs[torch.arange(len(b_idx), device=self.dev), b_idx % A]

    def update(self, adv, ret):
        cfg, T = self.cfg, self.cfg.episode_length
        stats = {}
        for _ in range(cfg.num_sgd_steps):
            pick = torch.randperm(T * self.B, generator=self.gen, device=self.dev)[:cfg.batch_size]
            t_idx, b_idx = pick // self.B, pick % self.B
            logits, values = self._policy(self._observations(t_idx, b_idx))
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
