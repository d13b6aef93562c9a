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


def test_wasserstein_1d_equals_scipy():
    """the device-resident 1-D Wasserstein distance against scipy.stats.wasserstein_distance, which the reference
    calls on host copies (guide_dm_trainer.py:277-279): unequal sizes, ties, single elements, heavy tails."""
    import numpy as np
    import torch
    from scipy.stats import wasserstein_distance
    from cld_b200.metrics import wasserstein_1d
    rng = np.random.default_rng(5)
    cases = [(rng.normal(size=1000), rng.normal(0.3, 2.0, size=777)),
             (rng.integers(0, 5, size=300).astype(np.float64), rng.integers(2, 9, size=450).astype(np.float64)),       # many ties
             (np.array([1.5]), np.array([-2.0])),
             (np.array([0.0, 0.0, 0.0]), np.array([0.0])),
             (rng.standard_cauchy(size=2000).astype(np.float32), rng.standard_cauchy(size=64).astype(np.float32))]
    for u, v in cases:
        want = wasserstein_distance(u, v)
        got = wasserstein_1d(torch.tensor(u), torch.tensor(v)).item()
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (got, want)
    import pytest
    with pytest.raises(ValueError):
        wasserstein_1d(torch.tensor([]), torch.tensor([1.0]))
    with pytest.raises(ValueError):
        wasserstein_1d(torch.tensor([float("nan")]), torch.tensor([1.0]))


def test_realism_accumulator_equals_the_reference_test_epoch_arithmetic():
    """The statistics of test_step / on_test_epoch_end (guide_dm_trainer.py:219-296) restated with numpy + scipy exactly as the
    reference writes them, against cld_b200.metrics on two batches; `last_batch_only` reproduces the reference's re-created lists."""
    import numpy as np
    import torch
    from scipy.stats import wasserstein_distance
    from cld_b200.metrics import RealismAccumulator
    torch.manual_seed(11)
    dt = 0.1
    batches = [(torch.randn(24, 52, 6), torch.randn(24, 52, 6) * 1.3 + 0.1), (torch.randn(8, 52, 6) * 0.7, torch.randn(8, 52, 6))]
    stats = [{"offroad_failure_rate": 0.25, "collision_failure_rate": 0.5, "overall_failure_rate": 0.375},
             {"offroad_failure_rate": 0.75, "collision_failure_rate": 0.0, "overall_failure_rate": 0.375}]

    def reference(bs):
        outs = []
        for pred, gt in bs:
            long_acc_gt, long_acc_pred = gt[..., 4], pred[..., 4]
            lat_acc_gt, lat_acc_pred = gt[..., 2] * gt[..., 5], pred[..., 2] * pred[..., 5]
            jerk_gt = (long_acc_gt[:, 1:] - long_acc_gt[:, :-1]) / dt
            jerk_pred = (long_acc_pred[:, 1:] - long_acc_pred[:, :-1]) / dt
            outs.append([a.numpy().flatten() for a in (long_acc_gt, long_acc_pred, lat_acc_gt, lat_acc_pred, jerk_gt, jerk_pred)])
        cat = [np.concatenate([o[i] for o in outs]) for i in range(6)]
        wl, wa, wj = wasserstein_distance(cat[0], cat[1]), wasserstein_distance(cat[2], cat[3]), wasserstein_distance(cat[4], cat[5])
        return wl, wa, wj, (wl + wa + wj) / 3.0

    acc = RealismAccumulator(dt)
    for (pred, gt), st in zip(batches, stats):
        acc.add_batch(pred, gt, st)
    got = acc.compute()
    want = reference(batches)
    for k, w in zip(("wd_long", "wd_lat", "wd_jerk", "realism_deviation"), want):
        assert abs(got[k] - w) < 1e-9, (k, got[k], w)
    assert abs(got["avg_offroad_failure_rate"] - 0.5) < 1e-12 and abs(got["avg_collision_failure_rate"] - 0.25) < 1e-12
    last = RealismAccumulator(dt, last_batch_only=True)
    for (pred, gt), st in zip(batches, stats):
        last.add_batch(pred, gt, st)
    got_last, want_last = last.compute(), reference(batches[-1:])
    assert abs(got_last["realism_deviation"] - want_last[3]) < 1e-9 and got_last["avg_offroad_failure_rate"] == 0.75


@pytest.mark.gpu
def test_test_step_mirror_feeds_realism_statistics(models_cpu):
    """cld_b200.metrics.test_step (guide_dm_trainer.py:205-247 on the B200 path): its statistics equal the reference's arithmetic
    (numpy + scipy) applied to the trajectories the oracle decodes from the SAME sampled latents."""
    from scipy.stats import wasserstein_distance
    from cld_b200.engine import default_guidance
    from cld_b200.metrics import RealismAccumulator, test_step
    dm, vae, algo = models_cpu(10)
    dm = dm.cuda()
    vae.bind(dm)
    S, A = 2, 4
    aux, batch = make_scenes(S, A, seed=77, dense=True)
    bd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    ad = {k: v.cuda() for k, v in aux.items()}
    torch.manual_seed(3)
    gt = torch.randn(S * A, 52, 6)
    acc = RealismAccumulator(algo.step_time)
    torch.manual_seed(4)
    stats = test_step(dm, vae, bd, ad, gt, algo, acc, guidance=default_guidance())
    got = acc.compute()
    assert set(stats) == {"offroad_failure_rate", "collision_failure_rate", "overall_failure_rate"}
    # the same sampled latents through the oracle's decoder + rollout, then the reference's arithmetic
    torch.manual_seed(4)
    out = dm(bd, ad, algo, guidance=default_guidance())
    dec_sd = {k: v.detach().cpu() for k, v in vae.lstmvae.lstm_dec.state_dict().items()}
    traj, _ = O.decode_rollout(dec_sd, out["pred_traj"].cpu(), aux["cond_feat"], aux["curr_states"])
    scaled = vae.scale_traj(traj)
    la_p, la_g = scaled[..., 4], gt[..., 4]
    lat_p, lat_g = scaled[..., 2] * scaled[..., 5], gt[..., 2] * gt[..., 5]
    dt = algo.step_time
    jp, jg = (la_p[:, 1:] - la_p[:, :-1]) / dt, (la_g[:, 1:] - la_g[:, :-1]) / dt
    want = [wasserstein_distance(a.numpy().flatten(), b.numpy().flatten()) for a, b in ((la_g, la_p), (lat_g, lat_p), (jg, jp))]
    for k, w in zip(("wd_long", "wd_lat", "wd_jerk"), want):
        assert abs(got[k] - w) <= 1e-3 * max(1.0, abs(w)), (k, got[k], w)
    assert abs(got["realism_deviation"] - sum(want) / 3.0) <= 1e-3 * max(1.0, sum(want) / 3.0)
    # validation_step mirror (guide_dm_trainer.py:185-202): mean reward of the decoded samples = the critic drop-in on the same latents
    from cld_b200.critic import compute_reward
    from cld_b200.metrics import validation_step
    torch.manual_seed(4)
    r = validation_step(dm, vae, bd, ad, algo, guidance=default_guidance())
    traj_d = vae.convert_action_to_state_and_action(vae.lstmvae.lstm_dec(out["pred_traj"], ad["cond_feat"]), ad["curr_states"], descaled_output=True)
    want_r = compute_reward(dm, traj_d.reshape(S * A, 1, 52, 6), bd).mean()
    assert torch.isfinite(r) and abs(r.item() - want_r.item()) <= 1e-5 * max(1.0, abs(want_r.item()))


# ---------------------------------------------------------------------------------------------------- f-3: closed-loop driver
class _StraightPolicy:
    """Stub of policy.get_action: every agent keeps its speed along its heading (agent frame: +x)."""

    def __init__(self, T, dt):
        self.T, self.dt = T, dt

    def get_action(self, obs, num_action_samples=1, step_index=0, sampler="ddpm", aux_info=None, **kw):
        v = obs["curr_speed"]
        tt = torch.arange(1, self.T + 1, device=v.device).float() * self.dt
        pos = torch.stack([v[:, None] * tt[None, :], torch.zeros(v.shape[0], self.T, device=v.device)], -1)
        return {"positions": pos, "yaws": torch.zeros(v.shape[0], self.T, 1, device=v.device)}, \
            {"act_idx": torch.zeros(v.shape[0], dtype=torch.long), "guide_losses": {}}


def test_closed_loop_driver_and_synthetic_env_frames():
    """rollout.py:95-100 on the synthetic environment (CPU, stub policy): observations are agent-centric (the newest history point
    is the origin, heading 0), the drivable raster of an agent's frame agrees with the world road, a constant-velocity plan moves
    every agent along its own heading by v * dt per step, the loop replans every n_step_action steps until is_done."""
    from cld_b200.rollout import SyntheticEnv, closed_loop_rollout
    env = SyntheticEnv(2, 3, num_steps=25, n_step_action=10, seed=4, device="cpu")
    p0, yaw0, v0 = env.pos.clone(), env.yaw.clone(), env.speed.clone()
    obs = env.get_observation()
    B = 6
    assert obs["history_positions"].shape == (B, 31, 2) and obs["drivable_map"].shape == (B, 224, 224)
    assert obs["history_positions"][:, -1].abs().max() < 1e-4 and obs["history_yaws"][:, -1].abs().max() < 1e-6
    assert torch.allclose(obs["agent_from_world"] @ obs["world_from_agent"], torch.eye(3).expand(B, 3, 3), atol=1e-5)
    # raster pixel of the agent's own position: (col 56, row 112); its drivable flag = the world road test at the agent
    assert torch.equal(obs["drivable_map"][:, 112, 56], env.drivable_world(env.pos))
    # the others' futures, mapped back to the world, are the other agents' plans
    back = torch.einsum('bij,bstj->bsti', obs["world_from_agent"][:, :2, :2], obs["all_other_agents_future_positions"]) + \
        obs["world_from_agent"][:, None, None, :2, 2]
    assert torch.allclose(back[0, 0], env.plan_world[1], atol=1e-3) and torch.allclose(back[1, 0], env.plan_world[0], atol=1e-3)
    recs = closed_loop_rollout(env, _StraightPolicy(env.T, env.dt), lambda o: None)
    assert [r["steps_taken"] for r in recs] == [10, 10, 5] and env.is_done() and env.t == 25
    want = p0 + torch.stack([torch.cos(yaw0), torch.sin(yaw0)], 1) * (v0 * 25 * env.dt)[:, None]
    assert torch.allclose(env.pos, want, atol=1e-3) and torch.allclose(env.speed, v0, atol=1e-3) and torch.allclose(env.yaw, yaw0)
    assert torch.equal(env.pos[v0 == 0], p0[v0 == 0])                       # parked agents (zero action) stay where they are
    m = env.metrics()
    assert m["steps"] == 25 and 0.0 <= m["offroad_rate"] <= 1.0 and 0.0 <= m["collision_rate"] <= 1.0


@pytest.mark.gpu
def test_closed_loop_rollout_with_the_guided_sampler(models_cpu):
    """The same loop with GuidedDiffusionPolicy.get_action on the B200 sampler (bf16, guided, 4 samples per agent, scene-level
    choice): runs to the end, replans 3 times, is deterministic under a fixed device-RNG seed, parked agents stay parked."""
    from cld_b200.policy import GuidedDiffusionPolicy
    from cld_b200.rollout import SyntheticEnv, closed_loop_rollout
    dm, vae, algo = models_cpu(10, precision="bf16")
    dm = dm.cuda()
    vae = vae.cuda().bind(dm)
    pol = GuidedDiffusionPolicy(dm, vae, algo, guidance=dict(agent_collision=50.0, map_collision=1.0))
    torch.manual_seed(3)
    cond = torch.randn(8, 256, device="cuda")
    finals = []
    for _ in range(2):
        env = SyntheticEnv(2, 4, num_steps=30, n_step_action=10, seed=7)
        parked = env.speed == 0
        p0 = env.pos.clone()
        recs = closed_loop_rollout(env, pol, lambda o: {"cond_feat": cond, "curr_states": env.curr_states()}, num_action_samples=4,
                                   use_device_rng=True, seed=11)
        assert len(recs) == 3 and env.t == 30 and all(r["act_idx"].shape == (8,) for r in recs)
        assert set(recs[0]["guide_losses"]) == {"agent_collision", "map_collision"}
        assert torch.isfinite(env.pos).all() and torch.equal(env.pos[parked], p0[parked])
        # scene-level choice: the agents of a scene share the sample index
        assert all((r["act_idx"].view(2, 4) == r["act_idx"].view(2, 4)[:, :1]).all() for r in recs)
        finals.append(env.pos.clone())
    assert torch.equal(finals[0], finals[1])
