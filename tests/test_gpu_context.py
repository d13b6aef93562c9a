"""GPU (-m gpu): the context encoder through the C ABI against the oracle and the golden
produced by the REAL reference (tests/golden/context.npz).

Tolerances: the convolutions run in bf16 with fp32 accumulation (17 layers deep), BatchNorm / residual / head in fp32:
ResNet stages <= 1e-2 relative L2, cond_feat <= 1e-2 relative L2 (measured values are printed)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _build(gold, max_agents=64):
    import cld_oracle as O
    from cld_b200 import default_algo_config
    from cld_b200.context import ContextEncoder
    g = gold("context")
    shapes = {str(k): eval(str(s)) for k, s in zip(g["keys"], g["shapes"])}
    sd = O.synth_context_state(shapes)
    ce = ContextEncoder(4, default_algo_config(), {"image": (34, 224, 224)}, max_agents=max_agents)
    ce.load_state_dict(sd)
    return g, sd, ce.cuda()


def _golden_batch(g):
    return {"image": torch.from_numpy(g["image_x2"]).float() / 2, "history_positions": torch.from_numpy(g["history_positions"]),
            "history_yaws": torch.from_numpy(g["history_yaws"]), "curr_speed": torch.from_numpy(g["curr_speed"])}


def test_context_stages_vs_oracle(gold):
    import cld_oracle as O
    g, sd, ce = _build(gold)
    batch = _golden_batch(g)
    with torch.no_grad():
        taps = {}
        O.context_encode(sd, batch, taps)
    cb = {k: v.cuda() for k, v in batch.items()}
    for stage, name in enumerate(("stem", "layer1", "layer2", "layer3", "layer4")):
        out = ce(cb, tap_stage=stage)
        torch.cuda.synchronize()
        r = rel(out["tap"], taps[name])
        print("context stage %s: rel %.3e" % (name, r))
        assert r < 1e-2, (name, r)


def test_context_vs_reference_golden(gold):
    g, sd, ce = _build(gold)
    cb = {k: v.cuda() for k, v in _golden_batch(g).items()}
    out = ce(cb, want_map_feat=True)
    torch.cuda.synchronize()
    r_map = rel(out["map_feat"], torch.from_numpy(g["map_feat"]))
    r = rel(out["cond_feat"], torch.from_numpy(g["cond_feat"]))
    print("context: rel(map_feat) %.3e rel(cond_feat) %.3e" % (r_map, r))
    assert torch.equal(out["curr_states"].cpu(), torch.from_numpy(g["curr_states"]))
    assert r_map < 1e-2 and r < 1e-2, (r_map, r)
    assert ce.launch_count() > 0


def test_context_batch_invariance_and_chunking(gold):
    """An agent's feature does not depend on what else is in the batch, on the tile it lands in, or on the workspace
    chunking (B = 21 agents: ragged last GEMM tile in every layer; chunk of 8 agents -> three passes)."""
    import cld_oracle as O
    from cld_b200.synthetic import make_context_batch
    g, sd, ce = _build(gold)
    batch = make_context_batch(21, seed=5)
    cb = {k: v.cuda() for k, v in batch.items()}
    full = ce(cb)["cond_feat"].clone()
    sub = {k: v[7:12].contiguous() for k, v in cb.items()}
    part = ce(sub)["cond_feat"]
    torch.cuda.synchronize()
    assert torch.equal(full[7:12], part)
    with torch.no_grad():
        want = O.context_encode(sd, batch)["cond_feat"]
    r = rel(full, want)
    print("context B=21: rel(cond_feat) %.3e" % r)
    assert r < 1e-2
    os.environ["CLD_CTX_CHUNK"] = "8"
    try:
        g2, sd2, ce2 = _build(gold)
        chunked = ce2(cb)["cond_feat"]
        torch.cuda.synchronize()
    finally:
        del os.environ["CLD_CTX_CHUNK"]
    assert torch.equal(chunked, full)


def test_context_crosses_image_box_boundaries(gold):
    """B = 133 agents: more than one 128-image box in layer4 (1x1x128 tiles), ragged boxes of 2 / 8 / 32 images in layers 1-3;
    every agent still matches the oracle and its own single-agent result."""
    import cld_oracle as O
    from cld_b200.synthetic import make_context_batch
    g, sd, ce = _build(gold, max_agents=133)
    batch = make_context_batch(133, seed=17)
    cb = {k: v.cuda() for k, v in batch.items()}
    full = ce(cb)["cond_feat"]
    one = ce({k: v[130:131].contiguous() for k, v in cb.items()})["cond_feat"]
    torch.cuda.synchronize()
    assert torch.equal(full[130:131], one)
    with torch.no_grad():
        want = O.context_encode(sd, batch)["cond_feat"]
    per_agent = ((full.cpu().double() - want.double()).norm(dim=1) / want.double().norm(dim=1)).max().item()
    print("context B=133: worst per-agent rel(cond_feat) %.3e" % per_agent)
    assert per_agent < 1e-2


def test_context_requires_cuda_module(gold):
    from cld_b200 import default_algo_config
    from cld_b200.context import ContextEncoder
    ce = ContextEncoder(4, default_algo_config(), {"image": (34, 224, 224)})
    with pytest.raises(RuntimeError):
        ce({"image": torch.zeros(1, 34, 224, 224), "history_positions": torch.zeros(1, 31, 2),
            "history_yaws": torch.zeros(1, 31, 1), "curr_speed": torch.zeros(1)})


def test_raster_to_trajectories_chain(gold):
    """The reference's inference chain end to end on the device: VaeModel.context_encoder(batch) -> DmModel(batch, aux_info)
    (guide_dm_trainer.py:85-90 with vae_model.py:84): cond_feat / curr_states produced by the context kernels feed the
    sampler unchanged."""
    from cld_b200 import default_algo_config, make_scenes
    from cld_b200.dm_model import DmModel
    from cld_b200.synthetic import make_context_batch
    from cld_b200.vae import VaeModel
    algo = default_algo_config()
    torch.manual_seed(0)
    dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=10, precision="bf16").cuda()
    vae = VaeModel(algo, None, {"image": (34, 224, 224)}).cuda().bind(dm)
    assert any(k.startswith("context_encoder.map_encoder.encoder_heads.map_model.layer4.1.bn2") for k in vae.state_dict())
    S, A = 2, 4
    _, batch = make_scenes(S, A, seed=9, dense=True)
    batch.update(make_context_batch(S * A, seed=4))
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    aux = vae.context_encoder(batch)
    assert aux["cond_feat"].shape == (S * A, 256) and aux["curr_states"].shape == (S * A, 4)
    out = dm(batch, {"cond_feat": aux["cond_feat"], "curr_states": aux["curr_states"]}, algo, want_indicators=True,
             agents_per_scene=A, seed=5)
    torch.cuda.synchronize()
    assert out["traj"].shape == (S * A, 52, 6) and torch.isfinite(out["traj"]).all()


def test_rasterize_agents_vs_reference_golden(gold):
    """The fused history rasteriser against the REAL rasterize_agents (tests/golden/raster.npz): bit-exact image."""
    g, sd, ce = _build(gold)
    r = gold("raster")
    img = ce.rasterize_agents(torch.from_numpy(r["maps_x2"]).float().cuda() / 2, torch.from_numpy(r["agent_hist_pos"]).cuda(), None,
                              torch.from_numpy(r["agent_hist_mask"]).cuda(), torch.from_numpy(r["raster_from_agent"]).cuda())
    torch.cuda.synchronize()
    want = torch.from_numpy(r["image_x2"]).float() / 2
    assert torch.equal(img.cpu(), want)


def test_forward_history_equals_forward_on_the_rasterised_image(gold):
    """Rasterising inside the encoder gives the same cond_feat as encoding the image rasterize_agents returns (bit for bit: the
    bf16 raster is identical), for a ragged batch and against the oracle chain."""
    import cld_oracle as O
    from cld_b200.synthetic import make_history_batch
    g, sd, ce = _build(gold)
    hb = make_history_batch(11, num_neighbors=7, seed=5)
    B = 11
    batch = {"raster_from_agent": hb["raster_from_agent"], "history_positions": hb["agent_hist_pos"][:, 0],
             "history_yaws": torch.zeros(B, 31, 1), "curr_speed": torch.rand(B) * 9}
    cb = {k: v.cuda() for k, v in batch.items()}
    fused = ce.forward_history(cb, hb["maps"].cuda(), hb["agent_hist_pos"].cuda(), hb["agent_hist_mask"].cuda(), want_image=True)
    cb["image"] = fused["image"]
    plain = ce(cb)
    torch.cuda.synchronize()
    assert torch.equal(fused["cond_feat"], plain["cond_feat"])
    img = O.rasterize_agents(hb["maps"], hb["agent_hist_pos"], hb["agent_hist_mask"], hb["raster_from_agent"])
    assert torch.equal(fused["image"].cpu(), img)
    with torch.no_grad():
        want = O.context_encode(sd, dict(batch, image=img))["cond_feat"]
    assert rel(fused["cond_feat"], want) < 1e-2


def test_context_kernel_variants_agree(gold):
    """The debug switches select other kernel variants for the same layers (single-CTA vs CTA-pair `cta_group::2`, plain vs
    grouped stages, N tile 128 vs 256): every variant must reproduce the default result (same bf16 operands and fp32
    accumulation; only the accumulation grouping differs)."""
    from cld_b200.synthetic import make_context_batch
    batch = {k: v.cuda() for k, v in make_context_batch(37, seed=23).items()}
    g, sd, ce = _build(gold)
    base = ce(batch)["cond_feat"].clone()
    torch.cuda.synchronize()
    for env in ({"CLD_CTX_PAIR": "0"}, {"CLD_CTX_PAIR": "2"}, {"CLD_CTX_GROUP": "0"}, {"CLD_CTX_NT": "128", "CLD_CTX_PAIR": "2"}):
        os.environ.update(env)
        try:
            _, _, ce2 = _build(gold)
            out = ce2(batch)["cond_feat"]
            torch.cuda.synchronize()
        finally:
            for k in env:
                del os.environ[k]
        r = rel(out, base)
        print("context variant %s: rel vs default %.3e" % (env, r))
        assert r < 2e-3, (env, r)


def test_context_permutation_equivariance_at_scale(gold):
    """Size-independent property at a large batch (700 agents: several head variants, many GEMM tiles, ragged image boxes):
    permuting the agents permutes the features, bit for bit, and a slice of the batch reproduces the same rows."""
    g, sd, ce = _build(gold, max_agents=700)
    B = 700
    gen = torch.Generator(device="cuda").manual_seed(3)
    img = (torch.rand(B, 34, 224, 224, device="cuda", generator=gen) < 0.03).float()
    batch = {"image": img, "history_positions": torch.randn(B, 31, 2, device="cuda", generator=gen),
             "history_yaws": torch.randn(B, 31, 1, device="cuda", generator=gen) * 0.1, "curr_speed": torch.rand(B, device="cuda", generator=gen) * 10}
    base = ce(batch)["cond_feat"].clone()
    perm = torch.randperm(B, device="cuda", generator=gen)
    out = ce({k: v[perm].contiguous() for k, v in batch.items()})["cond_feat"]
    torch.cuda.synchronize()
    assert torch.isfinite(base).all()
    assert torch.equal(out, base[perm])
    sub = ce({k: v[123:260].contiguous() for k, v in batch.items()})["cond_feat"]
    torch.cuda.synchronize()
    assert torch.equal(sub, base[123:260])
