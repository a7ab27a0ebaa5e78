"""Device-resident replay buffer (reference: algorithms/shared/replay.py:10-159).

The reference keeps 1e6 x (615*2+5) floats in pinned HOST memory, moves ~40 MB over PCIe per
environment step (ddpg.py:202-207, replay.py:94-143) and draws a CPU ``randperm(1e6)`` per
sample.  4.9 GB fits HBM trivially, so here everything -- storage, ring-buffer writes, sampling
without replacement -- stays on the device."""
import torch


class OffPolicyReplayBuffer(object):

    def __init__(self, state_dim, action_dim, max_size=int(1e6), device='cuda'):
        self.device = torch.device(device)
        self.max_size = int(max_size)
        self.ptr = 0
        self.size = 0
        kw = dict(dtype=torch.float32, device=self.device)
        self.state = torch.zeros((self.max_size, state_dim), **kw)
        self.action = torch.zeros((self.max_size, action_dim), **kw)
        self.next_state = torch.zeros((self.max_size, state_dim), **kw)
        self.reward = torch.zeros((self.max_size, 1), **kw)
        self.not_done = torch.zeros((self.max_size, 1), **kw)

    def add(self, state, action, next_state, reward, done):
        """Ring-buffer append of a batch of transitions (replay.py:56-87); tensors on any device."""
        n = state.shape[0]
        if n == 0:
            return
        ind = (torch.arange(n, device=self.device) + self.ptr) % self.max_size
        self.state[ind] = state.to(self.device, dtype=torch.float32)
        self.action[ind] = action.to(self.device, dtype=torch.float32)
        self.next_state[ind] = next_state.to(self.device, dtype=torch.float32)
        self.reward[ind] = reward.to(self.device, dtype=torch.float32).reshape(n, 1)
        self.not_done[ind] = 1. - done.to(self.device, dtype=torch.float32).reshape(n, 1)
        self.ptr = (self.ptr + n) % self.max_size
        self.size = min(self.size + n, self.max_size)

    def __len__(self):
        return self.size

    def sample(self, batch_size=4096, generator=None):
        """min(batch_size, size) transitions without replacement (replay.py:94-143)."""
        ind = torch.randperm(self.size, device=self.device, generator=generator)[:min(self.size, batch_size)]
        return (self.state.index_select(0, ind), self.action.index_select(0, ind),
                self.next_state.index_select(0, ind), self.reward.index_select(0, ind).squeeze(-1),
                self.not_done.index_select(0, ind).squeeze(-1))

    def clear_memory(self):
        self.ptr = 0
        self.size = 0
