"""SURVEY.md sec. 8 'next' rows f-2 / f-3, host side (CPU) and through the sampler (-m gpu)."""
import collections
import random

import numpy as np
import pytest
import torch

import cld_oracle as O
from cld_b200.synthetic import make_scenes


# ---------------------------------------------------------------------------------------------------- f-3
def test_choose_action_matches_reference_golden(gold):
    """cld_b200.policy.choose_action_from_guidance and the oracle's restatement against the REAL function's output."""
    from cld_b200.policy import choose_action_from_guidance
    g = gold("choose")
    for tag in ("scene", "agent"):
        losses, want = torch.tensor(g[tag + "_losses"]), torch.tensor(g[tag + "_idx"])
        assert torch.equal(choose_action_from_guidance(losses, losses.shape[0], tag == "scene"), want)
        assert torch.equal(O.choose_action_from_guidance(losses, losses.shape[0], tag == "scene"), want)
    # several scenes: each scene picks its own sample
    torch.manual_seed(0)
    l = torch.rand(3 * 4, 5, 2)
    idx = choose_action_from_guidance(l, 4, True)
    for s in range(3):
        assert (idx[s * 4:(s + 1) * 4] == torch.argmin(l[s * 4:(s + 1) * 4].sum(-1).sum(0))).all()


# ---------------------------------------------------------------------------------------------------- f-2
class _RefReplay:
    """The reference's ReplayBuffer (models/rl/criticmodel.py:147-187) restated: deque of per-row tuples."""

    def __init__(self, capacity, alpha):
        self.buffer, self.alpha, self.base, self.init = collections.deque(maxlen=capacity), alpha, 0.0, False

    def add(self, *ts):
        r = ts[3].mean().item()
        self.base = r if not self.init else self.alpha * self.base + (1 - self.alpha) * r
        self.init = True
        for i in range(ts[0].shape[0]):
            self.buffer.append(tuple(t[i] for t in ts))


def test_replay_buffer_matches_the_reference_semantics():
    from cld_b200.replay import ReplayBuffer, ppo_surrogate
    torch.manual_seed(1)
    mine, ref = ReplayBuffer(capacity=50, alpha=0.9), _RefReplay(50, 0.9)
    for n in (16, 16, 16, 16, 7, 60):                       # wraps the ring; the last add exceeds the capacity
        ts = (torch.randn(n, 52, 4), torch.randn(n, 52, 4), torch.randn(n), torch.randn(n), torch.randn(n, 256))
        mine.add(*ts)
        ref.add(*ts)
        assert len(mine) == len(ref.buffer)
        assert abs(mine.get_baseline() - ref.base) < 1e-12
        stored = torch.stack([t[2] for t in ref.buffer])    # log_p_old identifies a row
        x0, x1, lp, rw, cf = mine.sample(min(8, len(mine)))
        assert x0.shape[1:] == (52, 4) and cf.shape[1:] == (256,)
        for v in lp:
            assert (stored == v).any()
        assert len(set(lp.tolist())) == lp.numel()          # without replacement
    with pytest.raises(ValueError):
        mine.sample(51)
    tup = mine.sample_tuples(3)
    assert len(tup) == 3 and tup[0][0].shape == (52, 4)
    mine.clear()
    assert len(mine) == 0
    # the surrogate of guide_dm_trainer.py:158-168
    lp_new, lp_old, rew = torch.randn(9), torch.randn(9), torch.randn(9)
    adv = rew - 0.3
    ratios = torch.exp(lp_new - lp_old)
    want = -torch.min(ratios * adv, torch.clamp(ratios, 0.8, 1.2) * adv).mean()
    assert torch.allclose(ppo_surrogate(lp_new, lp_old, rew, 0.3), want)


@pytest.mark.gpu
def test_log_prob_and_device_replay_on_gpu(models_cpu):
    """DmModel.log_prob (models/dm/dm_model.py:165-174) in fp32 mode vs the oracle (1e-4), fed from a device-resident ReplayBuffer."""
    from cld_b200.replay import ReplayBuffer
    dm, vae, algo = models_cpu(16)
    dm = dm.cuda()
    torch.manual_seed(2)
    n = 24
    x1, x0, cond = torch.randn(n, 52, 4), torch.randn(n, 52, 4) * 0.3, torch.randn(n, 256)
    buf = ReplayBuffer(capacity=64)
    buf.add(x0.cuda(), x1.cuda(), torch.zeros(n).cuda(), torch.randn(n).cuda(), cond.cuda())
    bx0, bx1, _, _, bcond = buf.sample(n, generator=torch.Generator(device="cuda").manual_seed(0))
    assert bx0.is_cuda
    t = torch.zeros(n, dtype=torch.long)
    lp = dm.log_prob(bx1, bx0, {"cond_feat": bcond}, t.cuda())
    sd = {k: v.detach().cpu() for k, v in dm.model.state_dict().items()}
    sched = O.make_schedule(16)
    with torch.no_grad():
        eps = O.unet_forward(sd, bx1.cpu(), bcond.cpu(), t)
        mean = sched["x_t_cof"][0] * bx1.cpu() - sched["noise_cof"][0] * eps
        sigma = (0.5 * sched["posterior_log_variance_clipped"][0]).exp()
        want = torch.distributions.Normal(mean, sigma).log_prob(bx0.cpu()).mean(dim=(1, 2))
    r = ((lp.cpu().double() - want.double()).norm() / want.double().norm()).item()
    print("log_prob rel %.3e" % r)
    assert r < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("scene_level", [True, False])
def test_get_action_picks_the_sample_the_reference_rule_picks(models_cpu, scene_level):
    from cld_b200.engine import default_guidance
    from cld_b200.policy import GuidedDiffusionPolicy, choose_action_from_guidance
    dm, vae, algo = models_cpu(10, precision="bf16")
    dm = dm.cuda()
    vae.bind(dm)
    S, A, N = 3, 6, 5
    aux, batch = make_scenes(S, A, seed=17, dense=True)
    g = default_guidance() if scene_level else default_guidance(agent_collision=0.0, target_pos=1.0)
    if not scene_level:
        batch["target_pos"] = torch.stack([aux["curr_states"][:, 2] * 3.0, torch.zeros(S * A)], 1)
    pol = GuidedDiffusionPolicy(dm, vae, algo, guidance=g)
    cu = lambda d: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}       # noqa: E731
    torch.manual_seed(3)
    action, info = pol.get_action(cu(batch), num_action_samples=N, step_index=0, aux_info=cu(aux), use_device_rng=True, seed=5)
    assert action["positions"].shape == (S * A, 52, 2) and action["yaws"].shape == (S * A, 52, 1)
    assert info["action_samples"]["positions"].shape == (S * A, N, 52, 2)
    active = list(info["guide_losses"])
    assert ("agent_collision" in active) == scene_level
    stacked = torch.stack([info["guide_losses"][k] for k in active], dim=2).cpu()
    want = O.choose_action_from_guidance(stacked, A, scene_level)
    assert torch.equal(info["act_idx"].cpu(), want)
    if scene_level:
        assert all(len(set(want[s * A:(s + 1) * A].tolist())) == 1 for s in range(S))
    ar = torch.arange(S * A)
    assert torch.equal(action["positions"].cpu(), info["action_samples"]["positions"].cpu()[ar, want])
    stat = batch["curr_speed"].abs() < 0.5
    assert stat.any() and (action["positions"].cpu()[stat] == 0).all() and (info["action_samples"]["yaws"].cpu()[stat] == 0).all()
    assert algo.num_samp == 1                                    # restored
