"""GPU (-m gpu): the CUDA path (through the C ABI) against the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

import cld_oracle as O
from cld_b200.synthetic import make_scenes

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4          # north_star: denoiser outputs and trajectories within 1e-4 relative under fp32


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def gpu_models(models_cpu):
    cache = {}

    def build(n):
        if n not in cache:
            dm, vae, algo = models_cpu(n)
            dm = dm.cuda()
            vae.bind(dm)
            cache[n] = (dm, vae, algo)
        return cache[n]
    return build


def cpu_sd(m):
    return {k: v.detach().cpu() for k, v in m.state_dict().items()}


def test_library_loaded_and_arch():
    from cld_b200 import _lib
    assert _lib.lib.cld_version() == 100
    assert torch.cuda.get_device_capability()[0] == 10


def test_unet_fp32_vs_reference_golden(gpu_models, gold):
    g = gold("unet")
    dm, _, _ = gpu_models(10)
    eps = dm.denoise(torch.tensor(g["x"]).cuda(), {"cond_feat": torch.tensor(g["cond"]).cuda()}, torch.tensor(g["t"]).cuda())
    assert rel(eps, g["eps"]) < FP32_TOL
    assert rel(eps, g["eps"]) < 2e-5


def test_unet_fp32_every_stage_vs_oracle(gpu_models, gold):
    g = gold("unet")
    dm, _, _ = gpu_models(10)
    x, cond, t = torch.tensor(g["x"]), torch.tensor(g["cond"]), torch.tensor(g["t"])
    taps = {}
    with torch.no_grad():
        O.unet_forward(cpu_sd(dm.model), x, cond, t, taps=taps)
    names = ["downs.0.0", "downs.0.1", "downs.0.2", "downs.1.0", "downs.1.1", "downs.1.2", "downs.2.0", "downs.2.1",
             "mid_block1", "mid_block2", "ups.0.0", "ups.0.1", "ups.0.2", "ups.1.0", "ups.1.1", "ups.1.2", "final_conv.0"]
    eng = dm.engine(x.shape[0])
    for i, nm in enumerate(names):
        _, dbg = eng.unet_forward(x.cuda(), cond.cuda(), t.cuda(), debug_stage=i)
        want = taps[nm].reshape(x.shape[0], -1)
        assert dbg.shape == want.shape, nm
        assert rel(dbg, want) < 2e-5, nm


def test_unet_fp32_ragged_rows_and_T104(models_cpu, gold):
    # R not a multiple of any tile, and the long-horizon config (T=104)
    g = gold("unet")
    from cld_b200 import default_algo_config
    from cld_b200.dm_model import DmModel
    algo = default_algo_config(horizon=104)
    torch.manual_seed(0)
    dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=10).cuda()
    eps = dm.denoise(torch.tensor(g["x104"]).cuda(), {"cond_feat": torch.tensor(g["cond"][:3]).cuda()},
                     torch.tensor(g["t"][:3]).cuda())
    assert rel(eps, g["eps104"]) < 2e-5


def test_posterior_step_vs_oracle(gpu_models):
    dm, _, _ = gpu_models(100)
    eng = dm.engine(64)
    sch = O.make_schedule(100)
    torch.manual_seed(1)
    x, eps, nz = torch.randn(37, 52, 4), torch.randn(37, 52, 4), torch.randn(37, 52, 4)
    for i in (99, 50, 1, 0):
        mean, sig = O.ddpm_mean_sigma(sch, x, eps, i)
        want = mean + (0.0 if i == 0 else 1.0) * sig * nz
        got, gmean = eng.posterior_step(x.cuda(), eps.cuda(), nz.cuda(), i, want_mean=True)
        assert rel(got, want) < 1e-6 and rel(gmean, mean) < 1e-6
    for i, nxt in ((98, 96), (2, 0), (0, -1)):
        want = O.ddim_next(sch, x, eps, i, nxt)
        got = eng.posterior_step(x.cuda(), eps.cuda(), None, i, nxt, sampler="ddim")
        assert rel(got, want) < 1e-6


def test_cfg0_sampler_vs_reference_golden(gpu_models, gold):
    """BASELINE config 0: 1 scene x 16 agents, 10-step DDPM, same x_init / noise tensors / weights."""
    g = gold("cfg0_sample")
    dm, vae, algo = gpu_models(10)
    out = dm({"history_positions": torch.zeros(16, 31, 2)},
             {"cond_feat": torch.tensor(g["cond"]).cuda(), "curr_states": torch.tensor(g["curr"]).cuda()}, algo,
             noise=torch.tensor(g["noises"]).cuda(), x_init=torch.tensor(g["x_init"]).cuda())
    assert rel(out["pred_traj"], g["pred_traj"]) < FP32_TOL
    assert rel(out["x1"], g["x1"]) < FP32_TOL
    assert rel(out["log_prob_final"], g["log_prob_final"]) < 1e-6
    assert out["aux_info"]["cond_feat"].shape == (16, 256)
    # decode + rollout of the reference's own latent
    act = vae.lstmvae.lstm_dec(torch.tensor(g["pred_traj"]).cuda(), torch.tensor(g["cond"]).cuda())
    traj, act2 = vae.decode_to_trajectory(torch.tensor(g["pred_traj"]).cuda(), torch.tensor(g["cond"]).cuda(),
                                          torch.tensor(g["curr"]).cuda())
    assert rel(act, g["act"]) < 1e-5 and rel(act2, g["act"]) < 1e-5
    assert rel(traj, g["traj"]) < 1e-5
    t2 = vae.convert_action_to_state_and_action(torch.tensor(g["act"]).cuda(), torch.tensor(g["curr"]).cuda(),
                                                descaled_output=True)
    assert rel(t2, g["traj"]) < 1e-5


def test_strided_ddpm_vs_reference_golden(gpu_models, gold):
    g = gold("stride2_sample")
    dm, _, algo = gpu_models(100)
    dm.stride = 2
    try:
        out = dm({"history_positions": torch.zeros(4, 31, 2)}, {"cond_feat": torch.tensor(g["cond"]).cuda()}, algo,
                 noise=torch.tensor(g["noises"]).cuda(), x_init=torch.tensor(g["x_init"]).cuda())
    finally:
        dm.stride = 1
    assert out["x1"] is None
    assert rel(out["pred_traj"], g["pred_traj"]) < FP32_TOL


def test_ddim_sampler_vs_oracle(gpu_models):
    dm, _, algo = gpu_models(100)
    aux, _ = make_scenes(1, 8, seed=5)
    torch.manual_seed(2)
    x_init = torch.randn(8, 52, 4)
    dm.stride = 10
    try:
        out = dm({"history_positions": torch.zeros(8, 31, 2)}, {"cond_feat": aux["cond_feat"].cuda()}, algo,
                 x_init=x_init.cuda(), sampler="ddim")
    finally:
        dm.stride = 1
    with torch.no_grad():
        want = O.sample(cpu_sd(dm.model), O.make_schedule(100), aux["cond_feat"], x_init, None, 100, 10, "ddim")
    assert rel(out["pred_traj"], want["pred_traj"]) < FP32_TOL


def test_unicycle_vs_reference_golden(gpu_models, gold):
    g = gold("unicycle")
    dm, _, _ = gpu_models(10)
    st = dm.engine(32).unicycle(torch.tensor(g["curr"]).cuda(), torch.tensor(g["u"]).cuda())
    assert rel(st, g["state"]) < 1e-5


def test_indicators_bit_exact_vs_reference_golden(gpu_models, gold):
    g = gold("indicators")
    dm, _, _ = gpu_models(10)
    _, batch = make_scenes(2, 8, seed=31, dense=True)
    from cld_b200.critic import failure_rate_compute, indicators
    tr = torch.tensor(g["traj"]).cuda()
    off, coll, rew = indicators(dm, tr, batch)
    assert np.array_equal(off.cpu().numpy(), g["offroad"])
    assert np.array_equal(coll.cpu().numpy(), g["coll"])
    fr = failure_rate_compute(dm, tr, batch)
    for k in fr:
        assert abs(fr[k] - float(g[k])) < 1e-12
    want_rew = O.reward(torch.tensor(g["traj"]), batch)
    assert rel(rew, want_rew) < 1e-5


def test_indicators_with_samples_and_edge_maps(gpu_models):
    dm, vae, _ = gpu_models(10)
    S, A, N = 2, 5, 3
    aux, batch = make_scenes(S, A, seed=77, dense=True)
    batch["drivable_map"][0] = False                 # everything off-road
    batch["drivable_map"][1] = True                  # nothing off-road
    torch.manual_seed(8)
    u = torch.randn(S * A * N, 52, 2) * torch.tensor([4.0, 0.8])
    curr = aux["curr_states"].repeat_interleave(N, 0)
    st = O.unicycle_rollout(curr, u)
    st[4] = 1e4                                       # far outside the raster: clamps to the edge
    tr6 = torch.cat([st, u], -1)
    from cld_b200.critic import indicators
    off, coll, _ = indicators(dm, tr6.cuda(), batch, num_samp=N)
    rep = {k: (v.repeat_interleave(N, 0) if torch.is_tensor(v) and v.shape[0] == S * A else v) for k, v in batch.items()}
    woff, wcoll = O.indicators(tr6[..., :2], rep)
    assert torch.equal(off.cpu(), woff) and torch.equal(coll.cpu(), wcoll)
    assert off[:N].all() and not off[N:2 * N].any()


def test_guidance_step_vs_reference_golden(gpu_models, gold):
    g = gold("guidance")
    dm, vae, _ = gpu_models(10)
    S, A, N = int(g["S"]), int(g["A"]), int(g["N"])
    aux, batch = make_scenes(S, A, seed=int(g["seed"]), dense=True)
    from cld_b200.engine import default_guidance
    eng = dm.engine(S * A * N)
    scene = eng.make_scene(batch, S, A, N)
    z = torch.tensor(g["z"]).cuda()
    cond = aux["cond_feat"].repeat_interleave(N, 0).cuda()
    curr = aux["curr_states"].repeat_interleave(N, 0).cuda()
    z_out, grad, loss = eng.guidance_step(z, cond, curr, scene, default_guidance())
    assert rel(loss[0], g["loss_ac"].reshape(-1)) < 1e-4
    assert rel(loss[1], g["loss_mc"].reshape(-1)) < 1e-4
    # gradient level: oracle autograd (fp32, CPU)
    dec_sd = {k: v.detach().cpu() for k, v in vae.lstmvae.lstm_dec.state_dict().items()}
    g_or, _ = O.guidance_grad(dec_sd, torch.tensor(g["z"]), aux["cond_feat"], aux["curr_states"], batch, A, N)
    assert rel(grad, g_or) < 1e-3
    zero_agree = ((grad.cpu() == 0) == (g_or == 0)).float().mean().item()
    nz = g_or != 0
    sign_agree = (torch.sign(grad.cpu())[nz] == torch.sign(g_or)[nz]).float().mean().item()
    print("guidance: rel(grad)=%.3e sign agreement %.6f zero-set agreement %.6f" % (rel(grad, g_or), sign_agree, zero_agree))
    assert sign_agree > 0.999 and zero_agree > 0.999
    # the reference's own Adam step
    assert rel(z_out, g["z_out"]) < 5e-3
    frac_bad = ((z_out.cpu() - torch.tensor(g["z_out"])).abs() > 1e-3).float().mean().item()
    assert frac_bad < 1e-3


def test_guidance_target_pos_and_sgd_vs_oracle(gpu_models):
    dm, vae, _ = gpu_models(10)
    S, A, N = 2, 4, 2
    aux, batch = make_scenes(S, A, seed=91, dense=True)
    from cld_b200.engine import default_guidance
    cfg = default_guidance(target_pos=2.0, optimizer="sgd", lr=1.0)
    eng = dm.engine(S * A * N)
    scene = eng.make_scene(batch, S, A, N)
    torch.manual_seed(92)
    z = torch.randn(S * A * N, 52, 4)
    cond = aux["cond_feat"].repeat_interleave(N, 0)
    curr = aux["curr_states"].repeat_interleave(N, 0)
    z_out, grad, loss = eng.guidance_step(z.cuda(), cond.cuda(), curr.cuda(), scene, cfg)
    dec_sd = {k: v.detach().cpu() for k, v in vae.lstmvae.lstm_dec.state_dict().items()}
    ocfg = dict(O.DEFAULT_GUIDANCE, target_pos=2.0, optimizer="sgd", lr=1.0)
    g_or, per = O.guidance_grad(dec_sd, z, aux["cond_feat"], aux["curr_states"], batch, A, N, ocfg)
    assert rel(loss[2], torch.cat([p["target_pos"] for p in per]).reshape(-1)) < 1e-4
    assert rel(grad, g_or) < 1e-3
    assert rel(z_out, O.apply_guidance_update(z, g_or, ocfg)) < 1e-5


def test_guided_sampler_vs_oracle(gpu_models):
    """Guided strided-DDPM (cfg1-shaped, small): CUDA loop vs the composed oracle on identical noise."""
    dm, vae, algo = gpu_models(10)
    S, A, N = 2, 4, 1
    aux, batch = make_scenes(S, A, seed=55, dense=True)
    from cld_b200.engine import default_guidance
    torch.manual_seed(56)
    R = S * A * N
    x_init, noises = torch.randn(R, 52, 4), torch.randn(10, R, 52, 4)
    out = dm({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()},
             {k: v.cuda() for k, v in aux.items()}, algo, noise=noises.cuda(), x_init=x_init.cuda(),
             guidance=default_guidance(), want_indicators=True)
    dec_sd = {k: v.detach().cpu() for k, v in vae.lstmvae.lstm_dec.state_dict().items()}
    gd = dict(dec_sd=dec_sd, cond=aux["cond_feat"], curr=aux["curr_states"], batch=batch, A=A, N=N, cfg=O.DEFAULT_GUIDANCE)
    want = O.sample(cpu_sd(dm.model), O.make_schedule(10), aux["cond_feat"], x_init, noises, 10, 1, "ddpm", guidance=gd)
    r = rel(out["pred_traj"], want["pred_traj"])
    frac = ((out["pred_traj"].cpu() - want["pred_traj"]).abs() > 1e-2 * want["pred_traj"].abs().max()).float().mean().item()
    print("guided sampler: rel %.3e, fraction of elements off by >1%% of max: %.5f" % (r, frac))
    assert frac < 0.02
    wtraj, _ = O.decode_rollout(dec_sd, out["pred_traj"].cpu(), aux["cond_feat"], aux["curr_states"])
    assert rel(out["traj"], wtraj) < 1e-4
    woff, wcoll = O.indicators(out["traj"].cpu()[..., :2], batch)
    assert torch.equal(out["offroad"].cpu(), woff) and torch.equal(out["coll"].cpu(), wcoll)


def test_sampler_chunking_and_device_rng(gpu_models):
    """R > max_rows is processed in whole-scene chunks with identical results; Philox noise is N(0,1)."""
    from cld_b200 import default_algo_config
    from cld_b200.dm_model import DmModel
    dm, _, algo = gpu_models(10)
    torch.manual_seed(0)
    dm_small = DmModel(default_algo_config(), {"image": (34, 224, 224)}, n_timesteps=10, max_rows=8).cuda()
    aux, _ = make_scenes(3, 4, seed=3)
    torch.manual_seed(4)
    x_init, noises = torch.randn(12, 52, 4).cuda(), torch.randn(10, 12, 52, 4).cuda()
    eng_small = dm_small.engine(1)
    assert eng_small.max_rows == 8
    a = eng_small.sample(x_init, aux["cond_feat"].cuda(), noises=noises)
    b = dm.engine(12).sample(x_init, aux["cond_feat"].cuda(), noises=noises)
    assert torch.equal(a["x0"], b["x0"])
    eng = dm.engine(12)
    zero = torch.zeros(4096, 52, 4).cuda()
    c = eng.sample(zero[:12], aux["cond_feat"].cuda(), seed=1234)
    d = eng.sample(zero[:12], aux["cond_feat"].cuda(), seed=1234)
    e = eng.sample(zero[:12], aux["cond_feat"].cuda(), seed=1235)
    assert torch.equal(c["x0"], d["x0"]) and not torch.equal(c["x0"], e["x0"])


def test_error_paths(gpu_models):
    dm, _, _ = gpu_models(10)
    eng = dm.engine(16)
    with pytest.raises(RuntimeError):
        eng.posterior_step(torch.zeros(2, 52, 4).cuda(), torch.zeros(2, 52, 4).cuda(), None, 99)   # t out of range
    from cld_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(horizon=50)                       # horizon must be a multiple of 4
    e2 = Engine(n_timesteps=10, max_rows=4)
    with pytest.raises(RuntimeError):
        e2.unet_forward(torch.zeros(2, 52, 4).cuda(), torch.zeros(2, 256).cuda(), torch.zeros(2).long().cuda())  # no weights


def test_guidance_step_64_agents_vs_oracle(gpu_models):
    """cfg3-shaped scene (64 agents in one scene): agent-collision pairs across the whole scene, fp32 kernels vs oracle autograd."""
    dm, vae, _ = gpu_models(10)
    S, A, N = 1, 64, 1
    aux, batch = make_scenes(S, A, seed=61, dense=True)
    from cld_b200.engine import default_guidance
    eng = dm.engine(S * A * N)
    scene = eng.make_scene(batch, S, A, N)
    torch.manual_seed(62)
    z = torch.randn(S * A * N, 52, 4)
    z_out, grad, loss = eng.guidance_step(z.cuda(), aux["cond_feat"].cuda(), aux["curr_states"].cuda(), scene, default_guidance())
    dec_sd = {k: v.detach().cpu() for k, v in vae.lstmvae.lstm_dec.state_dict().items()}
    g_or, per = O.guidance_grad(dec_sd, z, aux["cond_feat"], aux["curr_states"], batch, A, N)
    assert rel(loss[0], torch.cat([p["agent_collision"] for p in per]).reshape(-1)) < 1e-4
    assert rel(loss[1], torch.cat([p["map_collision"] for p in per]).reshape(-1)) < 1e-4
    assert rel(grad, g_or) < 1e-3
    nz = g_or != 0
    assert (torch.sign(grad.cpu())[nz] == torch.sign(g_or)[nz]).float().mean().item() > 0.999


# ----------------------------------------------------------------------------------------------------
# sharding neutrality: Philox noise is indexed by the GLOBAL row id
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_shards_chunks_and_lanes_equal_the_unsharded_batch(models_cpu, precision):
    """Guided DDPM with in-kernel Philox noise on 8 scenes x 4 agents x 2 samples: (i) the two halves `shard_batch` gives to
    ranks 0 / 1 of a 2-rank job, each sampled with its global row offset, concatenated; (ii) an engine whose max_rows forces
    4 chunks; (iii) DmModel(lanes=2) -- all bit-identical to the single unsharded call, trajectories and indicators included."""
    from cld_b200.distributed import shard_batch, shard_scenes
    from cld_b200.engine import default_guidance
    S, A, N = 8, 4, 2
    aux, batch = make_scenes(S, A, seed=77, dense=True)
    # per-agent guidance inputs as well (target speeds, waypoints): their rows must follow the agents through shards, chunks and lanes
    from cld_b200.waypoints import TargetPosAtTime
    torch.manual_seed(9)
    batch["target_speed"] = (aux["curr_states"][:, 2:3] + torch.randn(S * A, 1)).clamp(min=0).repeat(1, 52)
    batch.update(TargetPosAtTime(batch["target_pos"], torch.randint(5, 52, (S * A,)), agents=torch.rand(S * A) < 0.7).scene_entries(batch, 52, A))
    guid = default_guidance(target_speed=2.0, waypoint=1.0)
    kw = dict(sampler="ddpm", guidance=guid, use_device_rng=True, seed=2024, want_indicators=True, agents_per_scene=A)

    def build(**k):
        dm, vae, algo = models_cpu(10, precision=precision, **k)
        algo.num_samp = N
        dm = dm.cuda()
        vae.bind(dm)
        return dm, algo
    dm, algo = build(max_rows=S * A * N)
    torch.manual_seed(5)
    x_init = torch.randn(S * A * N, 52, 4).cuda()
    cu = lambda d: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}       # noqa: E731
    full = dm(cu(batch), cu(aux), algo, x_init=x_init, **kw)
    assert torch.isfinite(full["pred_traj"]).all()
    # (i) two ranks
    parts = []
    for rank in range(2):
        s0, s1 = shard_scenes(S, 2, rank)
        r0, r1 = s0 * A * N, s1 * A * N
        parts.append(dm(cu(shard_batch(batch, A, 2, rank)), cu(shard_batch(aux, A, 2, rank)), algo, x_init=x_init[r0:r1],
                        row_offset=r0, **kw))
    for key in ("pred_traj", "traj", "offroad", "coll"):
        assert torch.equal(torch.cat([p[key] for p in parts]), full[key]), key
    # a different offset draws different noise
    other = dm(cu(shard_batch(batch, A, 2, 0)), cu(shard_batch(aux, A, 2, 0)), algo, x_init=x_init[:S * A * N // 2], row_offset=8, **kw)
    assert not torch.equal(other["pred_traj"], parts[0]["pred_traj"])
    # (ii) chunked by max_rows (2 scenes per chunk)
    dm_c, algo_c = build(max_rows=2 * A * N)
    eng = dm_c.engine(1)
    assert eng.max_rows == 2 * A * N
    scene = eng.make_scene(batch, S, A, N)
    o = eng.sample(x_init, aux["cond_feat"].cuda().repeat_interleave(N, 0), seed=2024, curr_rows=aux["curr_states"].cuda().repeat_interleave(N, 0),
                   scene=scene, guidance=guid, sampler="ddpm", want_indicators=True)
    assert torch.equal(o["x0"], full["pred_traj"]) and torch.equal(o["traj"], full["traj"]) and torch.equal(o["coll"], full["coll"])
    plain = dm(cu(batch), cu(aux), algo, x_init=x_init, **dict(kw, guidance=default_guidance()))
    assert not torch.equal(plain["pred_traj"], full["pred_traj"])            # the per-agent terms do act
    # (iii) two lanes
    dm_l, algo_l = build(max_rows=S * A * N, lanes=2)
    lan = dm_l(cu(batch), cu(aux), algo_l, x_init=x_init, **kw)
    for key in ("pred_traj", "traj", "offroad", "coll"):
        assert torch.equal(lan[key], full[key]), key
    # x_init drawn in-kernel: reproducible, offset-consistent, N(0,1)
    e2 = dm.engine(S * A * N)
    cond_rows = aux["cond_feat"].cuda().repeat_interleave(N, 0)
    a = e2.sample(None, cond_rows, seed=11, sampler="ddim", stride=10)
    b = e2.sample(None, cond_rows[16:48], seed=11, row_offset=16, sampler="ddim", stride=10)
    assert torch.equal(a["x0"][16:48], b["x0"])


def test_two_rank_nccl_job_equals_one_rank(tmp_path):
    """2 processes x 1 GPU each (NCCL): each samples its shard of 8 scenes, one all-gather; rank 0 compares with the unsharded
    single-GPU result.  Skipped on a box with fewer than 2 GPUs."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tools", "shard_check.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", script], capture_output=True, text=True, timeout=600)
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert out.returncode == 0 and "SHARD_CHECK_OK" in out.stdout


# ----------------------------------------------------------------------------------------------------
# guidance terms pinned to the REAL reference (tests/golden/guidance_ext.npz, guidance_t104.npz)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_guidance_terms_vs_reference_golden(models_cpu, gold, precision):
    """Rows a13 / f-4: TargetPosLoss (guidance_loss.py:672-712, min_target_time 0.3), TargetSpeedLoss (:219-254), AccLimitLoss (:1444-1468),
    SpeedLimitLoss (:1509-1538), each alone and all six terms together: per-row losses, the SGD-extracted gradient and the Adam update
    of the REAL PerturbationGuidance.perturb (tests/golden/guidance_ext.npz)."""
    from conftest import guidance_ext_case
    from cld_b200.engine import default_guidance
    g = gold("guidance_ext")
    dm, vae, _ = models_cpu(10, precision=precision)
    dm = dm.cuda()
    vae.bind(dm)
    S, A, N, aux, batch, z, cfgs = guidance_ext_case(g)
    eng = dm.engine(S * A * N)
    scene = eng.make_scene(batch, S, A, N)
    cond, curr = aux["cond_feat"].repeat_interleave(N, 0).cuda(), aux["curr_states"].repeat_interleave(N, 0).cuda()
    rows = {"agent_collision": 0, "map_collision": 1, "target_pos": 2, "target_speed": 3, "acc_limit": 4, "speed_limit": 5}
    # bf16 mode: the tensor-core LSTM decoder's actions differ by ~1e-4 relative, which thresholded terms (limits) amplify
    gtol, ztol, ltol = (1e-3, 3e-3, 1e-4) if precision == "fp32" else (5e-3, 1e-2, 1e-3)
    for tag, ocfg in cfgs.items():
        cfg = default_guidance(**{k: ocfg[k] for k in ("agent_collision", "map_collision", "target_pos", "target_speed", "acc_limit",
                                                       "acc_limit_value", "speed_limit", "speed_limit_value", "min_target_time") if k in ocfg})
        z_out, grad, loss = eng.guidance_step(z.cuda(), cond, curr, scene, cfg)
        for key, r in rows.items():
            name = "%s_loss_%s" % (tag, key)
            if name in g:
                assert rel(loss[r], g[name].reshape(-1)) < ltol, (tag, key)
        gref = torch.tensor(g[tag + "_grad_sgd"])
        big = gref.abs() > 1e-4 * gref.abs().max()
        rg = rel(grad.cpu()[big], gref[big])
        agree = (torch.sign(grad.cpu())[big] == torch.sign(gref)[big]).float().mean().item()
        rz = rel(z_out, g[tag + "_z_out"])
        print("%s %-12s rel(grad) %.2e sign agreement %.5f rel(z') %.2e" % (precision, tag, rg, agree, rz))
        assert rg < gtol and agree > 0.999 and rz < ztol, (tag, rg, agree, rz)


def test_guidance_t104_64_agents_8_samples_vs_reference_golden(models_cpu, gold):
    """cfg2 / cfg3 shapes: one scene of 64 agents x 8 samples, horizon 104, agent + map collision, against the real reference."""
    from conftest import guidance_big_case
    from cld_b200 import default_algo_config
    from cld_b200.dm_model import DmModel
    from cld_b200.engine import default_guidance
    from cld_b200.vae import VaeModel
    g = gold("guidance_t104")
    S, A, N, T, aux, batch, z = guidance_big_case(g)
    algo = default_algo_config(num_samp=N)
    algo.horizon = T
    torch.manual_seed(0)
    dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=10, max_rows=S * A * N).cuda()
    vae = VaeModel(algo).bind(dm)
    eng = dm.engine(S * A * N)
    scene = eng.make_scene(batch, S, A, N)
    cond, curr = aux["cond_feat"].repeat_interleave(N, 0).cuda(), aux["curr_states"].repeat_interleave(N, 0).cuda()
    z_out, grad, loss = eng.guidance_step(z.cuda(), cond, curr, scene, default_guidance())
    rows = torch.tensor(g["big_rows"]).long()
    assert rel(loss[0], g["big_loss_agent_collision"].reshape(-1)) < 1e-4
    assert rel(loss[1], g["big_loss_map_collision"].reshape(-1)) < 1e-4
    assert rel(grad.cpu()[rows[:96]], g["big_grad_rows"]) < 1e-3
    assert rel(grad.cpu().flatten(1).double().abs().sum(1), g["big_grad_rowabs"]) < 1e-3
    assert rel(z_out.cpu()[:64], g["big_z_out_head"]) < 3e-3


@pytest.mark.parametrize("packed", [False, True])
def test_map_collision_screen_edge_cases_vs_oracle(gpu_models, models_cpu, packed, monkeypatch):
    """The screened map-collision term (pixel-box test on the bit-packed map + work list + nearest on-road point among the per-grid-row
    candidates) on maps that stress it: a raster whose width is no multiple of 8, thin roads, checker patterns and pixel noise (many
    partially overlapping steps), agents that start at / beyond the raster border (clamped look-ups), one footprint larger than the
    screen's 56-pixel box limit (full path) and a stationary agent.  Checked (a) against an engine that searches exhaustively
    (CLD_MAP_EXHAUSTIVE=1): same minimum, argmin and tie count, so equal up to the order of the warp's partial sums; (b) against the
    oracle's autograd: losses everywhere, gradients on the maps where exact distance ties are not the rule (on a checker / noise map
    an off-road point has several nearest on-road points at the same grid distance and the pick is rounding noise on both sides)."""
    from cld_b200.engine import default_guidance
    from cld_b200.synthetic import pack_drivable_map
    dm, vae, _ = gpu_models(10)
    S, A, N = 3, 4, 2
    aux, batch = make_scenes(S, A, seed=311, dense=True)
    B = S * A
    H, W = 90, 100
    g = torch.Generator().manual_seed(7)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    dmap = torch.zeros(B, H, W, dtype=torch.bool)
    for b in range(B):
        kind = b % 4
        if kind == 0:
            dmap[b] = (yy - 45).abs() <= 3 + b                                    # thin horizontal road
        elif kind == 1:
            dmap[b] = ((yy // 5 + xx // 7) % 2 == 0)                              # checker pattern: road edges everywhere
        elif kind == 2:
            dmap[b] = (xx >= 97) | (yy <= 1) | ((xx - 50).abs() < 6)               # drivable stripes at the raster border
        else:
            dmap[b] = torch.rand(H, W, generator=g) > 0.5                         # pixel noise
    batch["drivable_map"] = dmap
    batch["raster_from_agent"] = torch.tensor([[2., 0., 20.], [0., 2., 45.], [0., 0., 1.]]).repeat(B, 1, 1)
    ext = batch["extent"].clone()
    ext[1, 0], ext[1, 1] = 36.0, 3.0                                              # 72 pixels long: beyond the screen's box limit
    batch["extent"] = ext
    curr = aux["curr_states"].clone()
    curr[2, 0], curr[2, 1] = -12.0, -30.0                                         # starts outside the raster (top-left)
    curr[3, 0] = 41.0                                                             # starts beyond the right border
    curr[5, 2] = 0.0
    batch["curr_speed"] = batch["curr_speed"].clone()
    batch["curr_speed"][5] = 0.0                                                  # stationary: no map term
    aux = dict(aux, curr_states=curr)
    gb = dict(batch)
    if packed:
        gb["drivable_map_bits"] = pack_drivable_map(dmap)
        gb["drivable_map_width"] = W
    cfg = default_guidance(agent_collision=0.0, map_collision=1.0, optimizer="sgd", lr=1.0)
    torch.manual_seed(312)
    z = torch.randn(B * N, 52, 4)
    cond = aux["cond_feat"].repeat_interleave(N, 0)
    curr_rows = curr.repeat_interleave(N, 0)
    eng = dm.engine(B * N)
    scene = eng.make_scene(gb, S, A, N)
    z_out, grad, loss = eng.guidance_step(z.cuda(), cond.cuda(), curr_rows.cuda(), scene, cfg)
    # (a) exhaustive search (the switch is read when the engine is created)
    monkeypatch.setenv("CLD_MAP_EXHAUSTIVE", "1")
    dm2, vae2, _ = models_cpu(10)
    dm2 = dm2.cuda()
    vae2.bind(dm2)
    eng2 = dm2.engine(B * N)
    _, grad2, loss2 = eng2.guidance_step(z.cuda(), cond.cuda(), curr_rows.cuda(), eng2.make_scene(gb, S, A, N), cfg)
    monkeypatch.delenv("CLD_MAP_EXHAUSTIVE")
    assert rel(grad, grad2) < 1e-5 and rel(loss[1], loss2[1]) < 1e-6
    assert torch.equal(grad == 0, grad2 == 0)
    # (b) the oracle
    dec_sd = {k: v.detach().cpu() for k, v in vae.lstmvae.lstm_dec.state_dict().items()}
    ocfg = dict(O.DEFAULT_GUIDANCE, agent_collision=0.0, map_collision=1.0, optimizer="sgd", lr=1.0)
    g_or, per = O.guidance_grad(dec_sd, z, aux["cond_feat"], curr, batch, A, N, ocfg)
    want_loss = torch.cat([p["map_collision"] for p in per]).reshape(-1)
    assert want_loss.abs().sum() > 0 and g_or.abs().sum() > 0                     # the case is not vacuous
    assert rel(loss[1], want_loss) < 1e-4
    zero_agree = ((grad.cpu() == 0) == (g_or == 0)).float().mean().item()
    assert zero_agree > 0.999
    rows = torch.tensor([b % 4 in (0, 2) for b in range(B)]).repeat_interleave(N)
    assert g_or[rows].abs().sum() > 0
    assert rel(grad.cpu()[rows], g_or[rows]) < 1e-3
    print("map screen edge cases: rel(grad) vs exhaustive %.2e, vs oracle on tie-free maps %.2e, on all maps %.2e" % (
        rel(grad, grad2), rel(grad.cpu()[rows], g_or[rows]), rel(grad, g_or)))
    # the sampler's path runs the same kernels
    out = eng.sample(z.cuda(), cond.cuda(), noises=None, curr_rows=curr_rows.cuda(), scene=scene, guidance=cfg, sampler="ddim")
    assert torch.isfinite(out["x0"]).all()


def test_ragged_scenes_equal_per_scene_calls(gpu_models):
    """A batch whose scenes hold 4, 2, 6 and 4 agents (trajdata batches are ragged; the reference masks scenes block-diagonally):
    the guided sampler + indicators on the whole batch equal the same scenes sampled one call per scene, row for row."""
    from cld_b200.critic import failure_rate_compute, indicators
    from cld_b200.engine import default_guidance
    dm, vae, algo = gpu_models(10)
    sizes = [4, 2, 6, 4]
    B = sum(sizes)
    aux, batch = make_scenes(1, B, seed=801, dense=True)                       # one pool of 16 agents, re-partitioned into 4 scenes
    batch["scene_index"] = torch.cat([torch.full((c,), 10 + i) for i, c in enumerate(sizes)])
    others = batch["all_other_agents_future_positions"]
    torch.manual_seed(802)
    x_init, noises = torch.randn(B, 52, 4), torch.randn(10, B, 52, 4)
    bd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    ad = {k: v.cuda() for k, v in aux.items()}
    out = dm(bd, ad, algo, noise=noises.cuda(), x_init=x_init.cuda(), guidance=default_guidance(), want_indicators=True)
    pos = 0
    for c in sizes:
        sl = slice(pos, pos + c)
        sb = {k: (v[sl].cuda() if (torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B) else v) for k, v in batch.items()}
        sa = {k: v[sl].cuda() for k, v in aux.items()}
        one = dm(sb, sa, algo, noise=noises[:, sl].cuda(), x_init=x_init[sl].cuda(), guidance=default_guidance(), want_indicators=True)
        for k in ("pred_traj", "traj", "offroad", "coll"):
            assert torch.equal(out[k][sl], one[k]), (k, c)
        pos += c
    # the critic drop-ins take the ragged batch as well
    off, coll, _ = indicators(dm, out["traj"], bd, 1)
    assert torch.equal(off, out["offroad"]) and torch.equal(coll, out["coll"])
    stats = failure_rate_compute(dm, out["traj"], bd)
    assert 0.0 <= stats["overall_failure_rate"] <= 1.0
    assert others.shape[0] == B


def test_waypoint_guidance_terms_vs_oracle(gpu_models):
    """TargetPosAtTimeLoss / GlobalTargetPosAtTimeLoss / GlobalTargetPosLoss through cld_guidance_step (analytic
    gradient through the decoder) against the oracle's autograd; the oracle's formulas and cld_b200.waypoints' host logic are pinned to
    the real reference classes by tests/test_oracle.py::test_waypoint_terms_vs_reference_golden.  2 scenes x 4 agents x 2 samples,
    every branch (none / at-time / final-distance hinge / progress hinge / TargetPosLoss) and an agent that has arrived."""
    from cld_b200.waypoints import GlobalTargetPos, GlobalTargetPosAtTime, TargetPosAtTime
    dm, vae, algo = gpu_models(10)
    S, A, N, T = 2, 4, 2, 52
    aux, batch = make_scenes(S, A, seed=41, dense=True)
    B = S * A
    wfa = batch["world_from_agent"]
    batch["agent_from_world"] = torch.linalg.inv(wfa)
    hist = torch.zeros(B, 31, 8)
    hist[:, :, 0] = -(torch.arange(30, -1, -1).float() * 0.1)[None, :] * aux["curr_states"][:, 2:3]
    batch["agent_hist"] = hist
    rng = torch.tensor([6.0, 15.0, 40.0, 90.0, 25.0, 0.3, 60.0, 12.0])
    lat = torch.tensor([1.0, -2.0, 3.0, -4.0, 2.0, 0.1, -1.0, 0.5])
    local = torch.stack([rng, lat], 1)
    world = torch.einsum('bij,bj->bi', wfa[:, :2, :2], local) + wfa[:, :2, 2]
    urg = torch.tensor([0.0, 0.3, 0.5, 0.8, 0.2, 0.1, 0.4, 0.6])
    pref = torch.tensor([1.5, 2.0, 3.0, 4.0, 2.5, 1.0, 2.0, 3.5])
    gat = GlobalTargetPosAtTime(world, torch.tensor([20, 40, 80, 150, 3, 56, 30, 200]), urg, pref, target_tolerance=2.0)
    gat.update(5)
    terms = {"at_time": TargetPosAtTime(local, torch.tensor([3, 10, 25, 51, 40, 0, 7, 30]), agents=torch.tensor([1, 1, 0, 1, 1, 1, 1, 0]).bool()),
             "global_at_time": gat,
             "global": GlobalTargetPos(world, urg, pref, min_progress_dist=0.5, target_tolerance=2.0)}
    torch.manual_seed(12)
    z = torch.randn(B * N, T, 4)
    rep = lambda v: v.repeat_interleave(N, dim=0)        # noqa: E731
    eng = dm.engine(B * N)
    dec_sd = cpu_sd(vae.lstmvae.lstm_dec)
    seen = set()
    for tag, term in terms.items():
        ent = term.scene_entries(batch, T, A)
        seen |= set(ent["wp_mode"].tolist())
        b2 = dict(batch, **ent)
        g = dict(O.DEFAULT_GUIDANCE, agent_collision=0.0, map_collision=0.0, waypoint=2.0, optimizer="sgd", lr=1.0)
        want_grad, per = O.guidance_grad(dec_sd, z, aux["cond_feat"], aux["curr_states"], b2, A, N, g)
        assert want_grad.abs().sum() > 0
        scene = eng.make_scene(b2, S, A, N)
        z_out, grad, loss = eng.guidance_step(z.cuda(), rep(aux["cond_feat"]).cuda(), rep(aux["curr_states"]).cuda(), scene, g)
        want_loss = torch.cat([p["waypoint"] for p in per]).reshape(-1)
        assert rel(loss[6], want_loss) < 1e-4, tag
        assert rel(grad, want_grad) < 1e-3, (tag, rel(grad, want_grad))
        assert rel(z_out, z - want_grad) < 1e-4
        # rows of agents without a waypoint carry no gradient
        idle = rep(ent["wp_mode"] == 0)
        assert grad.cpu()[idle].abs().max() == 0 if idle.any() else True
    assert seen == {0, 1, 2, 3, 4}
    assert terms["global"].have_reached.tolist() == [False] * 5 + [True] + [False] * 2
