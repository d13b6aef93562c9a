"""Pin the oracle against the REAL reference and write tests/golden/*.npz  (container-only).

Run:  python oracle/make_golden.py          (needs /root/reference; prints every comparison)
The .npz files hold seeded inputs + the REFERENCE's outputs (not the oracle's); tests/test_oracle.py
re-checks the oracle against them anywhere, tests/test_gpu_*.py check the CUDA path against them on
the B200.  Weights are not stored (17 MB): they are regenerated from `torch.manual_seed(0)` through
cld_b200's parameter containers, whose construction order mirrors the reference so the init is
bit-identical; weight checksums are stored to detect drift.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE]
import cld_oracle as O          # noqa: E402
import ref_harness as RH        # noqa: E402
from cld_b200.synthetic import make_scenes   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
torch.set_num_threads(8)


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def check(name, a, b, tol):
    r = rel(a.double(), b.double())
    print("%-40s rel=%.3e  max|d|=%.3e  %s" % (name, r, (a - b).abs().max().item(), "OK" if r <= tol else "FAIL"))
    assert r <= tol, name


def sd_np(sd):
    return {k: v.detach().numpy() for k, v in sd.items()}


def main():
    RH.install()
    from tbsim.utils.guidance_loss import PerturbationGuidance
    out = {}
    # ---------------- schedule --------------------------------------------------------------
    for n in (10, 16, 100):
        dm, vae, algo = RH.build_models(n_timesteps=n)
        sch = O.make_schedule(n)
        for k, v in sch.items():
            assert torch.equal(v, getattr(dm, k)), (n, k)
        print("schedule n=%d: 14 buffers bit-identical" % n)
    dm, vae, algo = RH.build_models(n_timesteps=10)
    unet_sd = {k: v.detach() for k, v in dm.model.state_dict().items()}
    dec_sd = {k: v.detach() for k, v in vae.lstmvae.lstm_dec.state_dict().items()}
    wsum = {
        'unet_sum': float(sum(v.double().sum() for v in unet_sd.values())),
        'unet_abs': float(sum(v.double().abs().sum() for v in unet_sd.values())),
        'dec_sum': float(sum(v.double().sum() for v in dec_sd.values())),
        'dec_abs': float(sum(v.double().abs().sum() for v in dec_sd.values())),
    }
    print("weight checksums", wsum)

    # ---------------- unet forward -----------------------------------------------------------
    torch.manual_seed(11)
    R = 6
    x = torch.randn(R, 52, 4)
    cond = torch.randn(R, 256)
    t = torch.tensor([0, 1, 3, 5, 8, 9])
    with torch.no_grad():
        eps_ref = dm.model(x, {'cond_feat': cond}, t)
        eps_or = O.unet_forward(unet_sd, x, cond, t)
    check("unet_forward T=52", eps_or, eps_ref, 2e-6)
    x104 = torch.randn(3, 104, 4)
    with torch.no_grad():
        e104_ref = dm.model(x104, {'cond_feat': cond[:3]}, t[:3])
        e104_or = O.unet_forward(unet_sd, x104, cond[:3], t[:3])
    check("unet_forward T=104", e104_or, e104_ref, 2e-6)
    np.savez_compressed(os.path.join(GOLD, "unet.npz"), x=x.numpy(), cond=cond.numpy(), t=t.numpy(),
                        eps=eps_ref.numpy(), x104=x104.numpy(), eps104=e104_ref.numpy(), **wsum)

    # ---------------- cfg0 sampler: 1 scene x 16 agents, DDPM-10 (Appendix A recipe) ------------
    torch.manual_seed(123)
    cond16 = torch.randn(16, 256)
    curr16 = torch.cat([torch.zeros(16, 2), torch.rand(16, 1) * 10, torch.zeros(16, 1)], 1)
    torch.manual_seed(7)
    with torch.no_grad():
        o = dm({'history_positions': torch.zeros(16, 31, 2)}, {'cond_feat': cond16, 'curr_states': curr16}, algo)
    # replay the same RNG stream to capture x_init and the per-step noises the reference drew
    torch.manual_seed(7)
    x_init = torch.randn(16, 1, 52, 4).reshape(16, 52, 4)
    # the reference draws randn_like at EVERY step (also t=0 where it is multiplied by 0)
    sch = O.make_schedule(10)
    steps = O.step_indices(10, 1)
    xs = x_init.clone()
    noises = []
    with torch.no_grad():
        for i in steps:
            tt = torch.full((16,), i, dtype=torch.long)
            eps = dm.model(xs, {'cond_feat': cond16}, tt)
            mean, sig = O.ddpm_mean_sigma(sch, xs, eps, i)
            nz = torch.randn_like(mean)
            noises.append(nz)
            xs = mean + (0.0 if i == 0 else 1.0) * sig * nz
    noises = torch.stack(noises)
    check("replay == reference DmModel.forward", xs, o['pred_traj'], 1e-6)
    with torch.no_grad():
        so = O.sample(unet_sd, sch, cond16, x_init, noises, 10, 1, 'ddpm')
    check("oracle sample cfg0 pred_traj", so['pred_traj'], o['pred_traj'], 1e-5)
    check("oracle sample cfg0 x1", so['x1'], o['x1'], 1e-5)
    check("oracle sample cfg0 log_prob_final", so['log_prob_final'], o['log_prob_final'], 1e-6)
    with torch.no_grad():
        act_ref = vae.lstmvae.lstm_dec(o['pred_traj'], cond16)
        traj_ref = vae.convert_action_to_state_and_action(act_ref, curr16, descaled_output=True)
        traj_or, act_or = O.decode_rollout(dec_sd, o['pred_traj'], cond16, curr16)
        traj_scaled_ref = vae.convert_action_to_state_and_action(act_ref, curr16)
    check("lstm_decode", act_or, act_ref, 1e-5)
    check("decode_rollout traj", traj_or, traj_ref, 1e-5)
    check("scale_traj", O.scale_traj(traj_or), traj_scaled_ref, 1e-5)
    np.savez_compressed(os.path.join(GOLD, "cfg0_sample.npz"), cond=cond16.numpy(), curr=curr16.numpy(),
                        x_init=x_init.numpy(), noises=noises.numpy(), pred_traj=o['pred_traj'].numpy(),
                        x1=o['x1'].numpy(), log_prob_final=o['log_prob_final'].numpy(),
                        act=act_ref.numpy(), traj=traj_ref.numpy(), **wsum)

    # ---------------- strided DDPM (n=100, stride 2) through the reference ------------------------
    dm100, _, _ = RH.build_models(n_timesteps=100)
    assert all(torch.equal(a, b) for a, b in zip(dm100.model.state_dict().values(), unet_sd.values()))
    dm100.stride = 2
    torch.manual_seed(5)
    with torch.no_grad():
        o2 = dm100({'history_positions': torch.zeros(4, 31, 2)}, {'cond_feat': cond16[:4], 'curr_states': curr16[:4]}, algo)
    torch.manual_seed(5)
    xi2 = torch.randn(4, 1, 52, 4).reshape(4, 52, 4)
    sch100 = O.make_schedule(100)
    xs = xi2.clone(); nz2 = []
    with torch.no_grad():
        for i in O.step_indices(100, 2):
            eps = dm100.model(xs, {'cond_feat': cond16[:4]}, torch.full((4,), i, dtype=torch.long))
            mean, sig = O.ddpm_mean_sigma(sch100, xs, eps, i)
            n_ = torch.randn_like(mean); nz2.append(n_)
            xs = mean + (0.0 if i == 0 else 1.0) * sig * n_
    nz2 = torch.stack(nz2)
    check("replay stride2 == reference", xs, o2['pred_traj'], 1e-6)
    assert o2['x1'] is None
    with torch.no_grad():
        so2 = O.sample(unet_sd, sch100, cond16[:4], xi2, nz2, 100, 2, 'ddpm')
    check("oracle sample n=100 stride=2", so2['pred_traj'], o2['pred_traj'], 2e-5)
    np.savez_compressed(os.path.join(GOLD, "stride2_sample.npz"), cond=cond16[:4].numpy(), x_init=xi2.numpy(),
                        noises=nz2.numpy(), pred_traj=o2['pred_traj'].numpy(), **wsum)

    # ---------------- rollout with saturating actions vs the reference ------------------------------
    torch.manual_seed(21)
    u = torch.randn(32, 52, 2) * torch.tensor([6.0, 1.5])
    c0 = torch.cat([torch.randn(32, 2), torch.rand(32, 1) * 28 - 2, torch.randn(32, 1)], 1)
    from tbsim.models.diffuser_helpers import unicyle_forward_dynamics
    st_ref = unicyle_forward_dynamics(dm.dyn, c0, u, 0.1, mode='parallel')
    check("unicycle (saturating)", O.unicycle_rollout(c0, u), st_ref, 1e-5)
    np.savez_compressed(os.path.join(GOLD, "unicycle.npz"), u=u.numpy(), curr=c0.numpy(), state=st_ref.numpy())

    # ---------------- indicators / failure rates -----------------------------------------------------
    from models.rl.criticmodel import failure_rate_compute, compute_collision_reward
    aux, batch = make_scenes(2, 8, seed=31, dense=True)
    torch.manual_seed(32)
    ui = torch.randn(16, 52, 2) * torch.tensor([3.0, 0.6])
    with torch.no_grad():
        st = O.unicycle_rollout(aux['curr_states'], ui)
        tr = torch.cat([st, ui], -1)
        # make some rows hit another agent's future so the collision indicator is exercised
        oth = batch['all_other_agents_future_positions']
        tr[3, :, :2] = oth[3, 1] + 0.3
        tr[9, 10:20, :2] = oth[9, 4, 10:20] - 0.5
    fr_ref = failure_rate_compute(tr, batch)
    fr_or = O.failure_rates(tr[..., :2], batch)
    print("failure rates ref", fr_ref, "oracle", fr_or)
    assert all(abs(fr_ref[k] - fr_or[k]) < 1e-9 for k in fr_ref)
    offroad, coll = O.indicators(tr[..., :2], batch)
    cr = compute_collision_reward(tr[..., :2], batch)
    assert torch.equal(-cr.view(-1), coll)
    np.savez_compressed(os.path.join(GOLD, "indicators.npz"), traj=tr.numpy(), offroad=offroad.numpy(),
                        coll=coll.numpy(), **{k: np.float64(v) for k, v in fr_ref.items()})

    # ---------------- guidance: real PerturbationGuidance.perturb, one scene per call -----------------
    def ref_perturb(z_mean, aux_s, batch_s, N, cfg_list, opt):
        def transform(x_dec, data_batch, params, bsize=None, num_samp=1):
            cs = aux_s['curr_states'].unsqueeze(1).expand(bsize, num_samp, 4).reshape(bsize * num_samp, 4)
            return vae.convert_action_to_state_and_action(x_dec, cs, scaled_input=True, descaled_output=True)
        pg = PerturbationGuidance(transform, {})
        pg.set_guidance([cfg_list])
        condN = aux_s['cond_feat'].repeat_interleave(N, 0)
        xg = z_mean.clone().detach().requires_grad_()
        x_out, per = pg.perturb(xg, batch_s, opt, num_samp=N, decoder=lambda zz: vae.lstmvae.lstm_dec(zz, condN))
        return x_out.detach(), per

    cfg_list = [
        {'name': 'agent_collision', 'weight': 50.0, 'agents': None,
         'params': {'num_disks': 2, 'buffer_dist': 0.2, 'decay_rate': 0.9, 'excluded_agents': None}},
        {'name': 'map_collision', 'weight': 1.0, 'agents': None,
         'params': {'num_points_lw': (10, 10), 'decay_rate': 0.9}},
    ]
    S, A, N = 3, 6, 2
    aux, batch = make_scenes(S, A, seed=41, dense=True)
    torch.manual_seed(42)
    zg = torch.randn(S * A * N, 52, 4)
    z_ref, loss_ref = [], {'agent_collision': [], 'map_collision': []}
    grads_ref = []
    for s in range(S):
        aux_s = {k: v[s * A:(s + 1) * A] for k, v in aux.items()}
        batch_s = O.slice_scene(batch, s * A, (s + 1) * A)
        batch_s = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch_s.items()}
        zs = zg[s * A * N:(s + 1) * A * N]
        # SGD lr=1 exposes the raw gradient:  z' = z - g
        zo_sgd, _ = ref_perturb(zs, aux_s, batch_s, N, cfg_list, {'optimizer': 'sgd', 'lr': 1.0, 'grad_steps': 1, 'perturb_th': None})
        grads_ref.append(zs - zo_sgd)
        zo, per = ref_perturb(zs, aux_s, batch_s, N, cfg_list, {'optimizer': 'adam', 'lr': 0.3, 'grad_steps': 1, 'perturb_th': None})
        z_ref.append(zo)
        loss_ref['agent_collision'].append(per['agent_collision_scene_000_00'])
        loss_ref['map_collision'].append(per['map_collision_scene_000_01'])
    z_ref = torch.cat(z_ref); grads_ref = torch.cat(grads_ref)
    loss_ref = {k: torch.cat(v) for k, v in loss_ref.items()}
    g_or, per_or = O.guidance_grad(dec_sd, zg, aux['cond_feat'], aux['curr_states'], batch, A, N)
    z_or = O.apply_guidance_update(zg, g_or)
    lo_ac = torch.cat([p['agent_collision'] for p in per_or]); lo_mc = torch.cat([p['map_collision'] for p in per_or])
    print("guidance: |g|max %.3e  nonzero frac %.3f  losses ac %.4e mc %.4e" % (
        grads_ref.abs().max(), (grads_ref != 0).float().mean(), loss_ref['agent_collision'].sum(), loss_ref['map_collision'].sum()))
    check("guidance loss agent_collision", lo_ac, loss_ref['agent_collision'], 1e-5)
    check("guidance loss map_collision", lo_mc, loss_ref['map_collision'], 1e-5)
    # the SGD-extracted reference gradient carries fp32 rounding of (z - (z - g)); compare loosely
    big = grads_ref.abs() > 1e-4 * grads_ref.abs().max()
    check("guidance grad (sgd-extracted)", g_or[big], grads_ref[big], 5e-3)
    sign_agree = (torch.sign(g_or) == torch.sign(z_ref.sub(zg).neg())).float().mean().item()
    print("sign agreement oracle-grad vs reference adam step: %.6f" % sign_agree)
    check("guidance adam update z'", z_or, z_ref, 2e-3)
    np.savez_compressed(os.path.join(GOLD, "guidance.npz"), S=S, A=A, N=N, seed=41, z=zg.numpy(), z_out=z_ref.numpy(),
                        grad_sgd=grads_ref.numpy(), loss_ac=loss_ref['agent_collision'].numpy(),
                        loss_mc=loss_ref['map_collision'].numpy(), **wsum)
    guidance_ext_golden()
    choose_golden()
    context_golden()
    raster_golden()
    print("golden files written to", GOLD)


def guidance_ext_golden():
    """Row a13 (TargetPosLoss), the f-4 analytic terms (TargetSpeedLoss, AccLimitLoss, SpeedLimitLoss) and the big shape
    (T = 104, 64 agents, 8 samples) against the REAL PerturbationGuidance.perturb -> tests/golden/guidance_ext.npz,
    guidance_t104.npz.  Every term is pinned alone (loss + SGD-extracted gradient) and all together (Adam update)."""
    RH.install()
    from tbsim.utils.guidance_loss import PerturbationGuidance
    dm, vae, algo = RH.build_models(n_timesteps=10)
    dec_sd = {k: v.detach() for k, v in vae.lstmvae.lstm_dec.state_dict().items()}

    def ref_perturb(z_mean, aux_s, batch_s, N, cfg_list, opt):
        def transform(x_dec, data_batch, params, bsize=None, num_samp=1):
            cs = aux_s['curr_states'].unsqueeze(1).expand(bsize, num_samp, 4).reshape(bsize * num_samp, 4)
            return vae.convert_action_to_state_and_action(x_dec, cs, scaled_input=True, descaled_output=True)
        pg = PerturbationGuidance(transform, {})
        pg.set_guidance([cfg_list])
        condN = aux_s['cond_feat'].repeat_interleave(N, 0)
        xg = z_mean.clone().detach().requires_grad_()
        x_out, per = pg.perturb(xg, batch_s, opt, num_samp=N, decoder=lambda zz: vae.lstmvae.lstm_dec(zz, condN))
        return x_out.detach(), per

    SGD = {'optimizer': 'sgd', 'lr': 1.0, 'grad_steps': 1, 'perturb_th': None}
    ADAM = {'optimizer': 'adam', 'lr': 0.3, 'grad_steps': 1, 'perturb_th': None}

    def run_ref(S, A, N, T, aux, batch, zg, terms, sgd_lr):
        """terms: list of (oracle key, reference cfg dict builder(scene slice) )"""
        z_out, grads, losses = [], [], {}
        for s in range(S):
            aux_s = {k: v[s * A:(s + 1) * A] for k, v in aux.items()}
            batch_s = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in O.slice_scene(batch, s * A, (s + 1) * A).items()}
            cfg_list = [mk(batch_s) for _, mk in terms]
            zs = zg[s * A * N:(s + 1) * A * N]
            # z' = z - lr * g: a large lr lifts the gradient above the fp32 rounding of the subtraction
            zo_sgd, _ = ref_perturb(zs, aux_s, batch_s, N, cfg_list, dict(SGD, lr=sgd_lr))
            grads.append((zs - zo_sgd) / sgd_lr)
            zo, per = ref_perturb(zs, aux_s, batch_s, N, cfg_list, ADAM)
            z_out.append(zo)
            for i, (key, _) in enumerate(terms):
                name = cfg_list[i]['name']
                losses.setdefault(key, []).append(per['%s_scene_000_%02d' % (name, i)])
        return torch.cat(z_out), torch.cat(grads), {k: torch.cat(v) for k, v in losses.items()}

    def compare(tag, S, A, N, aux, batch, zg, terms, g_cfg, out, grad_tol=1e-3, sgd_lr=100.0):
        z_ref, g_ref, l_ref = run_ref(S, A, N, zg.shape[1], aux, batch, zg, terms, sgd_lr)
        g_or, per_or = O.guidance_grad(dec_sd, zg, aux['cond_feat'], aux['curr_states'], batch, A, N, g_cfg)
        z_or = O.apply_guidance_update(zg, g_or, g_cfg)
        for key, _ in terms:
            lo = torch.cat([p[key] for p in per_or])
            check("%s loss %s" % (tag, key), lo, l_ref[key], 2e-5)
            out["%s_loss_%s" % (tag, key)] = l_ref[key].numpy()
        big = g_ref.abs() > 1e-4 * g_ref.abs().max()
        print("%s: |g|max %.3e nonzero frac %.3f" % (tag, g_ref.abs().max(), (g_ref != 0).float().mean()))
        check("%s grad (sgd-extracted)" % tag, g_or[big], g_ref[big], grad_tol)
        agree = (torch.sign(g_or) == torch.sign(zg - z_ref)).float().mean().item()
        print("%s: sign agreement oracle grad vs reference adam step %.6f" % (tag, agree))
        check("%s adam update z'" % tag, z_or, z_ref, 3e-3)
        out["%s_grad_sgd" % tag] = g_ref.numpy()
        out["%s_z_out" % tag] = z_ref.numpy()

    # ---- T = 52: every term alone, then all together
    S, A, N, T = 2, 5, 2, 52
    aux, batch = make_scenes(S, A, horizon=T, seed=51, dense=True)
    g = torch.Generator().manual_seed(52)
    batch['target_pos'] = torch.stack([aux['curr_states'][:, 2] * 3.0 + 2.0, torch.rand(S * A, generator=g) * 6 - 3], 1)
    batch['target_speed'] = (aux['curr_states'][:, 2:3] + torch.linspace(0, 3, T)[None] * (torch.rand(S * A, 1, generator=g) - 0.3)).contiguous()
    torch.manual_seed(53)
    zg = torch.randn(S * A * N, T, 4)
    ACC_LIM, SPD_LIM, MIN_TT = 0.05, 5.0, 0.3
    mk = {
        'agent_collision': lambda b: {'name': 'agent_collision', 'weight': 50.0, 'agents': None,
                                      'params': {'num_disks': 2, 'buffer_dist': 0.2, 'decay_rate': 0.9, 'excluded_agents': None}},
        'map_collision': lambda b: {'name': 'map_collision', 'weight': 1.0, 'agents': None,
                                    'params': {'num_points_lw': (10, 10), 'decay_rate': 0.9}},
        'target_pos': lambda b: {'name': 'target_pos', 'weight': 2.0, 'agents': None,
                                 'params': {'target_pos': b['target_pos'].clone(), 'min_target_time': MIN_TT}},
        'target_speed': lambda b: {'name': 'target_speed', 'weight': 3.0, 'agents': None,
                                   'params': {'dt': 0.1, 'target_speed': b['target_speed'].numpy().copy(),
                                              'fut_valid': np.ones(tuple(b['target_speed'].shape), dtype=bool)}},
        'acc_limit': lambda b: {'name': 'acc_limit', 'weight': 1.5, 'agents': None, 'params': {'acc_limit': ACC_LIM}},
        'speed_limit': lambda b: {'name': 'speed_limit', 'weight': 4.0, 'agents': None, 'params': {'speed_limit': SPD_LIM}},
    }
    wts = {'target_pos': 2.0, 'target_speed': 3.0, 'acc_limit': 1.5, 'speed_limit': 4.0, 'agent_collision': 50.0, 'map_collision': 1.0}
    base = dict(O.DEFAULT_GUIDANCE, agent_collision=0.0, map_collision=0.0, min_target_time=MIN_TT, acc_limit_value=ACC_LIM,
                speed_limit_value=SPD_LIM)
    out = {'S': S, 'A': A, 'N': N, 'seed': 51, 'z': zg.numpy(), 'target_pos': batch['target_pos'].numpy(),
           'target_speed': batch['target_speed'].numpy(), 'acc_limit_value': ACC_LIM, 'speed_limit_value': SPD_LIM,
           'min_target_time': MIN_TT, 'weights': np.array([wts[k] for k in ('target_pos', 'target_speed', 'acc_limit', 'speed_limit')])}
    for key in ('target_pos', 'target_speed', 'acc_limit', 'speed_limit'):
        compare(key, S, A, N, aux, batch, zg, [(key, mk[key])], dict(base, **{key: wts[key]}), out)
    order = ['agent_collision', 'map_collision', 'target_pos', 'target_speed', 'acc_limit', 'speed_limit']
    compare('all', S, A, N, aux, batch, zg, [(k, mk[k]) for k in order], dict(base, **wts), out)
    np.savez_compressed(os.path.join(GOLD, "guidance_ext.npz"), **out)

    # ---- T = 104, one scene of 64 agents, 8 samples (cfg3 / cfg2 shapes): agent + map collision
    S, A, N, T = 1, 64, 8, 104
    aux, batch = make_scenes(S, A, horizon=T, seed=61, dense=True)
    torch.manual_seed(62)
    zg = torch.randn(S * A * N, T, 4)
    out = {'S': S, 'A': A, 'N': N, 'T': T, 'seed': 61, 'z_seed': 62, 'z_sum': float(zg.double().sum())}
    compare('big', S, A, N, aux, batch, zg, [(k, mk[k]) for k in ('agent_collision', 'map_collision')], dict(O.DEFAULT_GUIDANCE), out, sgd_lr=1e4)
    # store compactly: gradient as fp32 of the non-zero ROWS only; z' is reproducible from z and the gradient up to the Adam eps,
    # so only its first 64 rows are kept as a spot check
    gfull = out.pop('big_grad_sgd'); zfull = out.pop('big_z_out')
    rows = np.nonzero(np.abs(gfull).reshape(gfull.shape[0], -1).sum(1) > 0)[0]
    out['big_rows'] = rows.astype(np.int32)
    out['big_grad_rows'] = gfull[rows[:96]]
    out['big_grad_rowsum'] = gfull.reshape(gfull.shape[0], -1).astype(np.float64).sum(1)
    out['big_grad_rowabs'] = np.abs(gfull).reshape(gfull.shape[0], -1).astype(np.float64).sum(1)
    out['big_z_out_head'] = zfull[:64]
    np.savez_compressed(os.path.join(GOLD, "guidance_t104.npz"), **out)
    print("guidance ext goldens written")


def choose_golden():
    """f-3: the REAL choose_action_from_guidance (src/tbsim/utils/guidance_loss.py:22-65) on random per-sample guidance losses of one
    scene (the reference's agent-centric batch: B = 1, M agents) -> tests/golden/choose.npz."""
    RH.install()
    from tbsim.utils.guidance_loss import choose_action_from_guidance
    import types
    g = torch.Generator().manual_seed(5)
    M, N, T = 7, 6, 52
    preds = {"positions": torch.zeros(M, N, T, 2)}
    out = {}
    for tag, names in (("scene", ["agent_collision", "map_collision"]), ("agent", ["map_collision", "target_pos"])):
        losses = {("%s_scene_000_%02d" % (n, i)): torch.rand(M, N, generator=g) for i, n in enumerate(names)}
        losses[list(losses)[1]][2, 3] = float("nan")                      # nansum path
        cfgs = [[types.SimpleNamespace(name=n) for n in names]]
        idx = choose_action_from_guidance(preds, {}, cfgs, losses)
        mine = O.choose_action_from_guidance(torch.stack(list(losses.values()), dim=2), M, tag == "scene")
        assert torch.equal(idx, mine), tag
        out[tag + "_losses"] = torch.stack(list(losses.values()), dim=2).numpy()
        out[tag + "_idx"] = idx.numpy()
    np.savez_compressed(os.path.join(GOLD, "choose.npz"), **out)
    print("choose_action_from_guidance: oracle == reference; golden written")


def context_golden():
    """a14: the REAL ContextEncoder (models/context_utils.py:8-61) on 3 synthetic agents with the deterministic
    parameter set of O.synth_context_state -> tests/golden/context.npz (raster as int8 = 2 x value, reference outputs)."""
    RH.install()
    from cld_b200.synthetic import make_context_batch
    _, vae, _ = RH.build_models(n_timesteps=10)
    ce = vae.context_encoder.eval()
    shapes = {k: tuple(v.shape) for k, v in ce.state_dict().items()}
    sd = O.synth_context_state(shapes)
    ce.load_state_dict(sd)
    batch = make_context_batch(3, seed=321)
    with torch.no_grad():
        ref = ce(batch)
        taps = {}
        mine = O.context_encode(sd, batch, taps)
    check("context cond_feat (oracle vs reference)", mine['cond_feat'], ref['cond_feat'], 1e-6)
    check("context curr_states", mine['curr_states'], ref['curr_states'], 0.0)
    img2 = (batch['image'] * 2).round()
    assert torch.equal(img2 / 2, batch['image'])
    np.savez_compressed(os.path.join(GOLD, "context.npz"), seed=321, image_x2=img2.to(torch.int8).numpy(),
                        history_positions=batch['history_positions'].numpy(), history_yaws=batch['history_yaws'].numpy(),
                        curr_speed=batch['curr_speed'].numpy(), cond_feat=ref['cond_feat'].numpy(),
                        curr_states=ref['curr_states'].numpy(), map_feat=taps['map_feat'].numpy(),
                        layer_absmean=np.array([taps[k].abs().mean().item() for k in ('stem', 'layer1', 'layer2', 'layer3', 'layer4')]),
                        w_sum=float(sum(v.double().sum() for v in sd.values())),
                        keys=np.array(list(shapes.keys())), shapes=np.array([str(v) for v in shapes.values()]))
    print("context golden written")


def raster_golden():
    """History rasterisation: the REAL rasterize_agents (src/tbsim/utils/trajdata_utils.py:123-156) on synthetic histories ->
    tests/golden/raster.npz (image as int8 = 2 x value)."""
    RH.install()
    from tbsim.utils.trajdata_utils import rasterize_agents
    from cld_b200.synthetic import make_history_batch
    hb = make_history_batch(4, num_neighbors=5, seed=77)
    yaw = torch.zeros(*hb['agent_hist_pos'].shape[:3], 1)
    ref = rasterize_agents(hb['maps'], hb['agent_hist_pos'], yaw, hb['agent_hist_mask'], hb['raster_from_agent'], None)
    mine = O.rasterize_agents(hb['maps'], hb['agent_hist_pos'], hb['agent_hist_mask'], hb['raster_from_agent'])
    assert torch.equal(ref, mine), "oracle rasterize_agents differs from the reference"
    n_ego, n_oth = int((ref[:, :31] == 1).sum()), int((ref[:, :31] == -1).sum())
    print("rasterize_agents: oracle == reference (bit-equal); ego pixels %d, other-agent pixels %d" % (n_ego, n_oth))
    assert torch.equal((ref * 2).round() / 2, ref)
    np.savez_compressed(os.path.join(GOLD, "raster.npz"), seed=77, image_x2=(ref * 2).round().to(torch.int8).numpy(),
                        maps_x2=(hb['maps'] * 2).round().to(torch.int8).numpy(), agent_hist_pos=hb['agent_hist_pos'].numpy(),
                        agent_hist_mask=hb['agent_hist_mask'].numpy(), raster_from_agent=hb['raster_from_agent'].numpy(),
                        n_ego=n_ego, n_oth=n_oth)
    print("raster golden written")



def ppo_golden():
    """one minibatch of the REAL `ppo_update` arithmetic (guide_dm_trainer.py:150-172: the reference's own
    DmModel.log_prob, the surrogate lines, loss.backward(), torch.optim.Adam) and of DmModel.compute_losses' MSE
    -> tests/golden/ppo.npz: inputs, log-probs, losses, per-tensor gradient norms / sums, sampled gradient and update values."""
    RH.install()
    dm, vae, algo = RH.build_models(n_timesteps=16)
    names = [k for k, _ in dm.model.named_parameters()]
    sched = O.make_schedule(16)
    unet_sd = {k: v.detach().clone() for k, v in dm.model.state_dict().items()}
    torch.manual_seed(2025)
    R = 12
    x1, cond = torch.randn(R, 52, 4), torch.randn(R, 256)
    # The trainer calls log_prob at t = 0 (guide_dm_trainer.py:160), where posterior_log_variance_clipped = log(1e-20), i.e.
    # sigma = 1e-10: (x0 - mean)^2 / (2 sigma^2) turns fp32 rounding of the mean (1e-7) into log-probs of -1e5 .. -1e6, so at t = 0
    # NO two fp32 implementations (not even the reference run twice with different batch compositions) agree.  The arithmetic is
    # pinned at t = 1 .. 15, where sigma is 0.02 .. 0.3; the t = 0 behaviour is covered by the tests as a property (finite or not,
    # the same formula).
    t = torch.tensor([1, 2, 3, 4, 5, 6, 8, 10, 12, 13, 14, 15])
    with torch.no_grad():
        eps = dm.model(x1, {'cond_feat': cond}, t)
        mean = dm.x_t_cof[t].reshape(-1, 1, 1) * x1 - dm.noise_cof[t].reshape(-1, 1, 1) * eps
        sigma = (0.5 * dm.posterior_log_variance_clipped[t]).exp().reshape(-1, 1, 1)
    x0 = mean + sigma * torch.randn(R, 52, 4)
    with torch.no_grad():
        lp_now = dm.log_prob(x1, x0, {'cond_feat': cond}, t)
    # old log-probs spread around the current ones so that ratios fall inside, above and below the clip range
    log_p_old = lp_now + torch.tensor([0.0, 0.05, -0.05, 0.15, -0.15, 0.3, -0.3, 0.5, -0.5, 0.1, -0.1, 0.0])
    reward = torch.randn(R) * 2 - 1
    baseline = -0.8
    for p in dm.model.parameters():
        p.requires_grad_(True)
    opt = torch.optim.Adam(dm.model.parameters(), lr=1e-4, weight_decay=1e-5)
    # ---- the reference's lines (guide_dm_trainer.py:155-172)
    advantage = reward - baseline
    log_p_new = dm.log_prob(x1, x0, {'cond_feat': cond}, t=t)
    ratios = torch.exp(log_p_new - log_p_old)
    surr1 = ratios * advantage
    surr2 = torch.clamp(ratios, 1 - 0.2, 1 + 0.2) * advantage
    loss = -torch.min(surr1, surr2).mean()
    opt.zero_grad()
    loss.backward()
    g_ref = {k: p.grad.detach().clone() for k, p in dm.model.named_parameters()}
    before = {k: p.detach().clone() for k, p in dm.model.named_parameters()}
    opt.step()
    delta = {k: (p.detach() - before[k]) for k, p in dm.model.named_parameters()}
    # ---- oracle == reference
    loss_or, lp_or, g_or = O.ppo_grads(unet_sd, sched, x1, x0, cond, t, log_p_old, reward, baseline, 0.2)
    check("ppo log_p_new (oracle vs reference)", lp_or, log_p_new.detach(), 1e-6)
    check("ppo loss", loss_or.reshape(1), loss.detach().reshape(1), 1e-6)
    worst = 0.0
    for k in names:
        worst = max(worst, rel(g_or[k].double(), g_ref[k].double()))
    print("ppo parameter gradients: worst per-tensor rel %.3e over %d tensors" % (worst, len(names)))
    assert worst < 1e-4
    p1, _, _ = O.adam_update(before[names[7]], g_ref[names[7]], torch.zeros_like(before[names[7]]), torch.zeros_like(before[names[7]]),
                             1, 1e-4, weight_decay=1e-5)
    check("adam_update restatement vs torch.optim.Adam", (p1 - before[names[7]].double()).float(), delta[names[7]], 1e-3)
    # ---- DM training loss (compute_losses with fixed t / noise)
    dm2, _, _ = RH.build_models(n_timesteps=16)
    torch.manual_seed(77)
    z0, tq, nz = torch.randn(R, 52, 4), torch.randint(0, 16, (R,)), torch.randn(R, 52, 4)
    z_noisy = dm2.q_sample(x_0=z0, t=tq, noise=nz)
    mse = torch.nn.functional.mse_loss(nz, dm2.model(z_noisy, {'cond_feat': cond}, tq))
    mse.backward()
    gm_ref = {k: p.grad.detach().clone() for k, p in dm2.model.named_parameters()}
    mse_or, gm_or = O.mse_grads(unet_sd, sched, z0, cond, tq, nz)
    check("mse loss (oracle vs reference)", mse_or.reshape(1), mse.detach().reshape(1), 1e-6)
    worst = max(rel(gm_or[k].double(), gm_ref[k].double()) for k in names)
    print("mse parameter gradients: worst per-tensor rel %.3e" % worst)
    assert worst < 1e-4
    # ---- golden: norms / sums of every tensor, 48 sampled entries of every tensor (fixed generator)
    gen = torch.Generator().manual_seed(5)
    idx = {k: torch.randint(0, g_ref[k].numel(), (48,), generator=gen) for k in names}
    pack = lambda d, f: np.stack([f(d[k]) for k in names])
    np.savez_compressed(
        os.path.join(GOLD, "ppo.npz"), x1=x1.numpy(), x0=x0.numpy(), cond=cond.numpy(), t=t.numpy(), log_p_old=log_p_old.numpy(),
        reward=reward.numpy(), baseline=baseline, clip=0.2, lr=1e-4, weight_decay=1e-5, names=np.array(names),
        log_p_new=log_p_new.detach().numpy(), loss=float(loss), ratios=ratios.detach().numpy(),
        grad_norm=pack(g_ref, lambda v: v.double().norm().item()), grad_sum=pack(g_ref, lambda v: v.double().sum().item()),
        grad_idx=np.stack([idx[k].numpy() for k in names]),
        grad_samples=np.stack([g_ref[k].reshape(-1)[idx[k]].numpy() for k in names]),
        delta_samples=np.stack([delta[k].reshape(-1)[idx[k]].numpy() for k in names]),
        delta_norm=pack(delta, lambda v: v.double().norm().item()),
        z0=z0.numpy(), tq=tq.numpy(), nz=nz.numpy(), mse=float(mse),
        mse_grad_norm=pack(gm_ref, lambda v: v.double().norm().item()),
        mse_grad_samples=np.stack([gm_ref[k].reshape(-1)[idx[k]].numpy() for k in names]))
    print("ppo golden written: loss %.6f, ratios %s" % (float(loss), np.round(ratios.detach().numpy(), 3)))



def waypoint_golden():
    """the REAL TargetPosAtTimeLoss / GlobalTargetPosAtTimeLoss / GlobalTargetPosLoss (guidance_loss.py:630-670,
    930-1135) on one scene of 8 agents x 3 samples -> tests/golden/waypoint.npz (inputs, the reference's losses [B,N] and the
    gradient of their sum w.r.t. the trajectories); the oracle's plan + formula must agree."""
    RH.install()
    from tbsim.utils import guidance_loss as GL
    torch.manual_seed(31)
    B, N, T, dt = 8, 3, 52, 0.1
    v = torch.rand(B, 1, 1) * 8 + 1
    tt = torch.arange(1, T + 1).float() * dt
    x = torch.zeros(B, N, T, 6)
    x[..., 0] = v * tt + 0.3 * torch.randn(B, N, 1).cumsum(0) * tt
    x[..., 1] = 0.4 * torch.randn(B, N, 1) * tt + 0.05 * torch.randn(B, N, T).cumsum(-1)
    yaw = (torch.rand(B) * 2 - 1) * 3.0
    pos = (torch.rand(B, 2) * 2 - 1) * 40
    c, s_ = torch.cos(yaw), torch.sin(yaw)
    wfa = torch.zeros(B, 3, 3)
    wfa[:, 0, 0], wfa[:, 0, 1], wfa[:, 0, 2] = c, -s_, pos[:, 0]
    wfa[:, 1, 0], wfa[:, 1, 1], wfa[:, 1, 2] = s_, c, pos[:, 1]
    wfa[:, 2, 2] = 1
    afw = torch.linalg.inv(wfa)
    hist = torch.zeros(B, 31, 8)
    hist[:, :, 0] = -(torch.arange(30, -1, -1).float() * dt)[None, :] * v[:, 0]           # straight past along -x in the agent frame
    batch = {'agent_from_world': afw, 'world_from_agent': wfa, 'agent_hist': hist}
    # world targets: ahead of every agent at different ranges; agents 6, 7 are AT their target (reached within the tolerance)
    rng = torch.tensor([6.0, 15.0, 30.0, 60.0, 90.0, 12.0, 0.5, 1.0])
    lat = torch.tensor([1.0, -2.0, 3.0, -4.0, 2.0, 0.5, 0.2, -0.3])
    tgt_local = torch.stack([rng, lat], 1)
    tgt_world = O._tf_points(tgt_local[:, None], wfa)[:, 0]
    out = dict(x=x.numpy(), world_from_agent=wfa.numpy(), agent_from_world=afw.numpy(), agent_hist=hist.numpy(),
               target_world=tgt_world.numpy(), target_local=tgt_local.numpy(), T=T, dt=dt)
    cases = {}
    # (1) TargetPosAtTimeLoss: local targets, per-agent time steps
    t_at = torch.tensor([3, 10, 25, 51, 40, 0, 7, 30])
    cases['at_time'] = (GL.TargetPosAtTimeLoss(tgt_local, t_at), dict(kind='target_pos_at_time', target_pos=tgt_local, target_time=t_at))
    # (2) GlobalTargetPosAtTimeLoss at global_t = 5: exact (time within the horizon), progress (beyond), passed (negative), reached
    t_gl = torch.tensor([20, 40, 80, 150, 3, 56, 30, 200])
    urg = torch.tensor([0.0, 0.3, 0.5, 0.8, 0.2, 0.1, 0.4, 0.6])
    pref = torch.tensor([1.5, 2.0, 3.0, 4.0, 2.5, 1.0, 2.0, 3.5])
    l2 = GL.GlobalTargetPosAtTimeLoss(tgt_world.numpy(), t_gl.numpy(), urg.numpy(), pref.numpy(), dt=dt, target_tolerance=4.5, action_num=5)
    l2.global_t = 5
    cases['global_at_time'] = (l2, dict(kind='global_target_pos_at_time', target_pos=tgt_world, target_time=t_gl, global_t=5, urgency=urg,
                                        pref_speed=pref, target_tolerance=4.5, action_num=5))
    # (3) GlobalTargetPosLoss: within reach -> TargetPosLoss, else progress; reached agents masked
    l3 = GL.GlobalTargetPosLoss(tgt_world.numpy(), urg.numpy(), pref.numpy(), dt=dt, min_progress_dist=0.5, target_tolerance=4.5, action_num=5)
    cases['global'] = (l3, dict(kind='global_target_pos', target_pos=tgt_world, urgency=urg, pref_speed=pref, min_progress_dist=0.5,
                                target_tolerance=4.5, action_num=5))
    w_samp = torch.rand(B, N) + 0.5                            # a non-uniform weighting so that every entry's gradient is exercised
    for tag, (ref_loss, kw) in cases.items():
        xr = x.clone().requires_grad_(True)
        l_ref = ref_loss(xr, batch, agt_mask=None)
        (g_ref,) = torch.autograd.grad((l_ref * w_samp).sum(), xr)
        wp = O.waypoint_plan(T=T, dt=dt, agent_from_world=afw, world_from_agent=wfa, agent_hist=hist, **kw)
        xo = x.clone().requires_grad_(True)
        l_or = O.waypoint_loss(xo, wp)
        (g_or,) = torch.autograd.grad((l_or * w_samp).sum(), xo)
        check("waypoint %s: loss (oracle vs reference)" % tag, l_or.detach(), l_ref.detach(), 1e-6)
        check("waypoint %s: d/dx" % tag, g_or, g_ref, 1e-6)
        print("   modes", wp['mode'].tolist(), "reached", wp['have_reached'].long().tolist())
        if hasattr(ref_loss, 'have_reached_mask') and ref_loss.have_reached_mask is not None:
            assert torch.equal(ref_loss.have_reached_mask[:, 0], wp['have_reached'])
        out.update({tag + '_loss': l_ref.detach().numpy(), tag + '_grad': g_ref.numpy(), tag + '_mode': wp['mode'].numpy(),
                    tag + '_time': wp['time'].numpy(), tag + '_dist': wp['dist'].numpy(), tag + '_target': wp['target'].numpy()})
    out.update(w_samp=w_samp.numpy(), t_at=t_at.numpy(), t_gl=t_gl.numpy(), urgency=urg.numpy(), pref_speed=pref.numpy(), global_t=5,
               target_tolerance=4.5, action_num=5, min_progress_dist=0.5)
    np.savez_compressed(os.path.join(GOLD, "waypoint.npz"), **out)
    print("waypoint golden written")


if __name__ == "__main__":
    if "--only-choose" in sys.argv:
        choose_golden()
    elif "--only-guidance-ext" in sys.argv:
        guidance_ext_golden()
    elif "--only-raster" in sys.argv:
        raster_golden()
    elif "--only-context" in sys.argv:
        context_golden()
    elif "--only-ppo" in sys.argv:
        ppo_golden()
    elif "--only-waypoint" in sys.argv:
        waypoint_golden()
    else:
        main()
