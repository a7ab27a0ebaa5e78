"""SAC with automatic temperature: the update step and its data-parallel hook
(reference: algorithms/sac_auto.py:37-250, shared/offpolicy.py:61-232).

The update is library-call territory (cuBLAS GEMMs through torch autograd) -- what this module
adds for the B200 build is where the work lives and how it scales:
  * the networks are plain torch modules whose parameters ARE the tensors the tcgen05 inference
    actor packs its bf16 weights from, so a rollout sees new weights after one repack launch;
  * with ``torch.distributed`` initialised every optimiser step is preceded by ONE all-reduce of
    that optimiser's gradients (alpha: 1 float, actor: 10.9 MB, critic: 21.9 MB), which live in one
    flat buffer per optimiser and are reduced in place, asynchronously, overlapped with the next
    backward pass -- SURVEY.md section 8(e): replicas stay in lock-step, polyak targets stay local.
"""
import copy
import math

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

LOG_STD_MAX = 2
LOG_STD_MIN = -20


def make_fc_network(widths, input_size, output_size):
    """Linear+ReLU stack with a linear last layer; state_dict keys {0,2,4,..}.{weight,bias}
    (reference: algorithms/shared/utils.py:41-51)."""
    dims = [input_size] + [int(w) for w in widths] + [output_size]
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class TorchMaxEntropyActor(nn.Module):
    """Autograd twin of the inference actor (offpolicy.py:61-140)."""

    def __init__(self, state_dim, action_dim, widths):
        super().__init__()
        self.action_dim = action_dim
        self.layers = make_fc_network(widths, state_dim, action_dim * 2)

    def forward(self, state, probabilistic=1.0, eps=None):
        p = self.layers(state)
        mu, log_std = p[:, :self.action_dim], p[:, self.action_dim:]
        std = torch.exp(torch.clamp(log_std, LOG_STD_MIN, LOG_STD_MAX)) * probabilistic
        if eps is None:
            eps = torch.randn_like(mu)
        pi = mu + std * eps                                     # Normal(mu, std).rsample()
        logp = (-((pi - mu) ** 2) / (2 * std ** 2) - torch.log(std) - math.log(math.sqrt(2 * math.pi))).sum(-1)
        logp = logp - (2 * (np.log(2) - pi - F.softplus(-2 * pi))).sum(1)
        return torch.tanh(pi), logp


class TorchDoubleCritic(nn.Module):
    """offpolicy.py:183-232: two Q networks over concat(state, action)."""

    def __init__(self, state_dim, action_dim, widths):
        super().__init__()
        self.q1 = make_fc_network(widths, state_dim + action_dim, 1)
        self.q2 = make_fc_network(widths, state_dim + action_dim, 1)

    def forward(self, state, action):
        x = torch.cat([state, action], -1)
        return self.q1(x).squeeze(-1), self.q2(x).squeeze(-1)


class _FlatGrads(object):
    """The gradients of one optimiser as views into ONE flat buffer, so that the data-parallel reduction
    is a single in-place all-reduce of that buffer with nothing to flatten or to copy back.  ``launch``
    starts the all-reduce asynchronously (NCCL runs it on its own stream while autograd carries on with
    the next backward pass); ``wait`` makes the averaged gradients visible before ``optimizer.step()``."""

    def __init__(self, params):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros((n,), dtype=ref.dtype, device=ref.device)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()
        self.work = None

    def zero(self):
        self.flat.zero_()
        o = 0
        for p in self.params:          # an optimiser's zero_grad(set_to_none=True) elsewhere must not detach the views
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * o:
                p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def launch(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
            self.flat /= dist.get_world_size()


class SACAutoLearner(object):
    """Networks, optimisers and ``update`` of SACAuto (sac_auto.py:37-250)."""

    def __init__(self, input_size, action_size, hidden_dims, lr=3e-4, gamma=0.99, alpha=0.2, device='cuda'):
        widths = [int(w) for w in str(hidden_dims).split('-')]
        self.device = torch.device(device)
        self.gamma = gamma
        self.tau = 0.005
        self.actor = TorchMaxEntropyActor(input_size, action_size, widths).to(self.device)
        self.critic = TorchDoubleCritic(input_size, action_size, widths).to(self.device)
        self.target_actor = copy.deepcopy(self.actor)
        self.target_critic = copy.deepcopy(self.critic)
        self.target_entropy = -float(np.prod(action_size))
        self.log_alpha = torch.full((1,), float(np.log(alpha)), requires_grad=True, device=self.device)
        self.alpha_optimizer = torch.optim.Adam([self.log_alpha], lr=lr)
        self.actor_optimizer = torch.optim.Adam(self.actor.parameters(), lr=lr)
        self.critic_optimizer = torch.optim.Adam(self.critic.parameters(), lr=lr)
        self.g_alpha = _FlatGrads([self.log_alpha])
        self.g_actor = _FlatGrads(list(self.actor.parameters()))
        self.g_critic = _FlatGrads(list(self.critic.parameters()))
        self.total_it = 0

    def broadcast_parameters(self, src=0):
        """Identical initial weights on every rank."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            for p in list(self.actor.parameters()) + list(self.critic.parameters()) + [self.log_alpha]:
                dist.broadcast(p.data, src)
            self.target_actor.load_state_dict(self.actor.state_dict())
            self.target_critic.load_state_dict(self.critic.state_dict())

    def update(self, batch, eps=None, eps_next=None):
        """One SAC-auto update on (state, action, next_state, reward, not_done).
        ``eps`` / ``eps_next`` fix the policy's N(0,1) draws (tests)."""
        self.total_it += 1
        state, action, next_state, reward, not_done = batch
        pi, logp_pi = self.actor(state, 1.0, eps)
        alpha_loss = -(self.log_alpha * (logp_pi + self.target_entropy).detach()).mean()
        alpha = self.log_alpha.exp()
        q1, q2 = self.critic(state, pi)
        actor_loss = (alpha * logp_pi - torch.min(q1, q2)).mean()
        with torch.no_grad():
            next_action, logp_next = self.actor(next_state, 1.0, eps_next)
            tq1, tq2 = self.target_critic(next_state, next_action)
            backup = reward + self.gamma * not_done * (torch.min(tq1, tq2) - alpha * logp_next)
        cq1, cq2 = self.critic(state, action)
        critic_loss = F.mse_loss(cq1, backup) + F.mse_loss(cq2, backup)

        # sac_auto.py:220-232: three backward passes, three optimiser steps, in the reference's order.
        # Data-parallel: each optimiser's gradients are reduced by ONE all-reduce of a flat buffer.  The
        # temperature (one float; the actor loss deposits a stray gradient in it afterwards, which the
        # reference discards the same way) is reduced and stepped at once; the actor's and the critic's
        # reductions are asynchronous -- the critic's backward pass runs while the actor's 10.9 MB travel
        # -- and are waited for just before their steps.  (The actor loss also deposits gradients in the
        # critic; clearing the critic's buffer after the actor's backward pass is the reference's
        # zero_grad() order.)
        self.g_alpha.zero()
        alpha_loss.backward()
        self.g_alpha.launch()
        self.g_alpha.wait()
        self.alpha_optimizer.step()

        self.g_actor.zero()
        actor_loss.backward()
        self.g_actor.launch()

        self.g_critic.zero()
        critic_loss.backward()
        self.g_critic.launch()

        self.g_actor.wait()
        self.actor_optimizer.step()
        self.g_critic.wait()
        self.critic_optimizer.step()

        with torch.no_grad():
            for net, tgt in ((self.critic, self.target_critic), (self.actor, self.target_actor)):
                for p, tp in zip(net.parameters(), tgt.parameters()):
                    tp.data.copy_(self.tau * p.data + (1 - self.tau) * tp.data)
        return {'actor_loss': actor_loss.detach(), 'critic_loss': critic_loss.detach(),
                'alpha_loss': alpha_loss.detach(), 'alpha': alpha.detach()}
