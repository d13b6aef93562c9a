"""Device-resident drop-in for the reference's `ReplayBuffer` (models/rl/criticmodel.py:147-187) -- SURVEY.md sec. 8 f-2.

The reference appends one Python tuple per ROW after a `.detach().cpu()` of each of its five tensors and later re-stacks random
tuples and moves them back (`guide_dm_trainer.py:127-156`).  Here the buffer is five preallocated device tensors used as a ring:
`add` is five slice copies, `sample` one `randint` + five gathers; nothing leaves HBM.  Same methods, same running reward
baseline (`alpha`-EMA of the batch mean reward); `sample` returns the stacked batch (what the PPO loop builds from the tuples),
`sample_tuples` the reference's list of per-row tuples for code that still zips them.
"""
import torch


class ReplayBuffer:
    def __init__(self, capacity=10000, alpha=0.9, device=None):
        self.capacity = int(capacity)
        self.alpha = float(alpha)
        self.device = torch.device(device) if device is not None else None
        self.running_reward_baseline = 0.0
        self.has_init_baseline = False
        self._bufs = None
        self._size = 0
        self._head = 0

    def _alloc(self, tensors):
        dev = self.device or tensors[0].device
        self.device = dev
        self._bufs = [torch.empty((self.capacity,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for t in tensors]

    def add(self, x0, x1, log_p_old, reward, cond_feat_value):
        """x0, x1 [n,T,D]; log_p_old, reward [n]; cond_feat_value [n,C] (criticmodel.py:155-176)."""
        ts = [x0.detach(), x1.detach(), log_p_old.detach(), reward.detach(), cond_feat_value.detach()]
        r = float(reward.mean().item())
        if not self.has_init_baseline:
            self.running_reward_baseline, self.has_init_baseline = r, True
        else:
            self.running_reward_baseline = self.alpha * self.running_reward_baseline + (1 - self.alpha) * r
        if self._bufs is None:
            self._alloc(ts)
        n = ts[0].shape[0]
        if n >= self.capacity:                      # deque(maxlen) keeps the newest `capacity` rows
            for b, t in zip(self._bufs, ts):
                b.copy_(t[n - self.capacity:])
            self._size, self._head = self.capacity, 0
            return
        first = min(n, self.capacity - self._head)
        for b, t in zip(self._bufs, ts):
            b[self._head:self._head + first].copy_(t[:first])
            if n > first:
                b[:n - first].copy_(t[first:])
        self._head = (self._head + n) % self.capacity
        self._size = min(self.capacity, self._size + n)

    def get_baseline(self):
        return self.running_reward_baseline

    def _indices(self, batch_size, generator=None):
        if batch_size > self._size:
            raise ValueError("Sample larger than population or is negative")      # random.sample's error
        # without replacement, like random.sample
        return torch.randperm(self._size, device=self.device, generator=generator)[:batch_size]

    def sample(self, batch_size, generator=None, out=None):
        """-> (x0 [b,T,D], x1 [b,T,D], log_p_old [b], reward [b], cond_feat [b,C]) on the device; `out`: five preallocated tensors to
        gather into (the static inputs of a CUDA graph)."""
        idx = self._indices(batch_size, generator)
        if out is not None:
            for b, o in zip(self._bufs, out):
                torch.index_select(b, 0, idx, out=o)
            return tuple(out)
        return tuple(b.index_select(0, idx) for b in self._bufs)

    def sample_tuples(self, batch_size, generator=None):
        cols = self.sample(batch_size, generator)
        return [tuple(c[i] for c in cols) for i in range(batch_size)]

    def clear(self):
        self._size = self._head = 0

    def __len__(self):
        return self._size


def ppo_surrogate(log_p_new, log_p_old, reward, baseline, eps=0.2):
    """The clipped surrogate of guide_dm_trainer.py:158-168 (forward value; the denoiser's backward is not built)."""
    adv = reward - baseline
    ratios = torch.exp(log_p_new - log_p_old)
    return -torch.min(ratios * adv, torch.clamp(ratios, 1 - eps, 1 + eps) * adv).mean()
