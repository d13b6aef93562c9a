"""CPU: host-side drop-in contract and the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_layout_matches_reference(models_cpu):
    dm, vae, _ = models_cpu(100)
    sd = dm.state_dict()
    keys = list(sd.keys())
    assert len(keys) == 162                       # 14 schedule buffers + 148 model tensors
    assert keys[:3] == ["betas", "alphas_cumprod", "alphas_cumprod_prev"]
    assert keys[13] == "noise_cof"
    mk = [k for k in keys if k.startswith("model.")]
    assert len(mk) == 148
    assert mk[0] == "model.time_mlp.1.weight" and mk[-1] == "model.final_conv.1.bias"
    assert sd["model.downs.0.0.blocks.0.block.0.weight"].shape == (64, 4, 5)
    assert sd["model.ups.0.0.blocks.0.block.0.weight"].shape == (128, 512, 5)
    assert sd["model.ups.0.2.conv.weight"].shape == (128, 128, 4)         # ConvTranspose1d [Cin,Cout,4]
    assert sd["model.downs.1.2.conv.weight"].shape == (128, 128, 3)
    assert "model.downs.2.2.conv.weight" not in sd and "model.downs.0.1.residual_conv.weight" not in sd
    assert sum(v.numel() for k, v in sd.items() if k.startswith("model.")) == 4349284
    vk = list(vae.state_dict().keys())
    for k in ["lstmvae.lstm_dec.lstm.weight_ih_l0", "lstmvae.lstm_dec.cond2hidden.weight",
              "lstmvae.lstm_dec.hid2act.bias", "lstmvae.lstm_enc.lstm.weight_hh_l1", "lstmvae.mu.weight"]:
        assert k in vk
    assert sum(v.numel() for v in vae.lstmvae.state_dict().values()) == 136458


def test_checkpoint_roundtrip_with_dm_prefix(models_cpu):
    dm, _, _ = models_cpu(100)
    ckpt = {"state_dict": {"dm." + k: v.clone() for k, v in dm.state_dict().items()}}
    dm2, _, _ = models_cpu(100)
    with torch.no_grad():
        for p in dm2.parameters():
            p.zero_()
    stripped = {k[3:]: v for k, v in ckpt["state_dict"].items() if k.startswith("dm.")}
    missing, unexpected = dm2.load_state_dict(stripped, strict=False)
    assert not missing and not unexpected
    assert all(torch.equal(a, b) for a, b in zip(dm.state_dict().values(), dm2.state_dict().values()))


def test_schedule_buffers_bit_identical_to_oracle(models_cpu):
    import cld_oracle as O
    for n in (10, 16, 100):
        dm, _, _ = models_cpu(n)
        for k, v in O.make_schedule(n).items():
            assert torch.equal(getattr(dm, k), v), (n, k)
    assert dm.n_timesteps == 100 and dm.stride == 1


def test_config_base_semantics():
    from cld_b200 import default_algo_config
    c = default_algo_config(num_samp=4)
    assert c.vae.latent_size == 4 and c["vae"]["hidden_size"] == 64
    assert c.dynamics["max_steer"] == 0.5 and "horizon" in c
    with pytest.raises(KeyError):
        c["nope"]
    assert c.get("nope", 7) == 7 and c.num_samp == 4


def test_no_cpu_fallback(models_cpu):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    dm, vae, algo = models_cpu(10)
    with pytest.raises(RuntimeError):
        dm({"history_positions": torch.zeros(2, 31, 2)}, {"cond_feat": torch.zeros(2, 256)}, algo)
    with pytest.raises(RuntimeError):
        vae.lstmvae.lstm_dec(torch.zeros(2, 52, 4), torch.zeros(2, 256))


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cld_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cld_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 16
    so = os.path.join(ROOT, "controllable-latent-diffusion-for-traffic-simulation_b200", "libcld_b200.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(lib, name), name
    from cld_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared
    assert lib.cld_version() == 100


def test_synthetic_scene_shapes():
    from cld_b200 import make_scenes
    aux, b = make_scenes(3, 5, horizon=52, seed=1)
    assert aux["cond_feat"].shape == (15, 256) and aux["curr_states"].shape == (15, 4)
    assert b["drivable_map"].shape == (15, 224, 224) and b["drivable_map"].dtype == torch.bool
    assert b["all_other_agents_future_positions"].shape == (15, 4, 52, 2)
    assert b["scene_index"].tolist() == [0] * 5 + [1] * 5 + [2] * 5
    assert torch.equal(b["raster_from_agent"][0], torch.tensor([[2., 0., 56.], [0., 2., 112.], [0., 0., 1.]]))


def test_context_encoder_has_no_cpu_path():
    """cld_b200.ContextEncoder is a parameter container + C-ABI call: on a CPU module it must raise, not fall back."""
    import pytest
    import torch
    from cld_b200 import default_algo_config
    from cld_b200.context import ContextEncoder
    ce = ContextEncoder(4, default_algo_config(), {"image": (34, 224, 224)})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ce({"image": torch.zeros(1, 34, 224, 224), "history_positions": torch.zeros(1, 31, 2),
            "history_yaws": torch.zeros(1, 31, 1), "curr_speed": torch.zeros(1)})
    with pytest.raises(RuntimeError):
        ContextEncoder(4, default_algo_config(), {"image": (3, 224, 224)})


def test_ragged_or_interleaved_scenes_are_rejected():
    """every scene of a call must hold the same number of contiguous agents; anything else raises a clear error
    instead of silently mis-partitioning rows (the reference builds a block-diagonal mask, guidance_loss.py:493-503)."""
    from cld_b200.keys import agents_per_scene
    assert agents_per_scene(torch.tensor([0, 0, 0, 1, 1, 1]), 6) == 3
    assert agents_per_scene(torch.tensor([7, 7, 3, 3]), 4) == 2
    assert agents_per_scene(None, 5) == 5
    with pytest.raises(ValueError, match="same number of agents"):
        agents_per_scene(torch.tensor([0] * 4 + [1] * 2 + [2] * 6), 12)       # B % A == 0 but ragged
    with pytest.raises(ValueError, match="not contiguous"):
        agents_per_scene(torch.tensor([0, 1, 0, 1]), 4)
    with pytest.raises(ValueError):
        agents_per_scene(torch.tensor([0, 0, 1]), 4)


def test_ragged_batches_are_bucketed_by_scene_size():
    """Scenes of different sizes (the reference accepts them through a block-diagonal mask, guidance_loss.py:493-503) are sampled as
    one uniform sub-batch per scene size: the buckets partition the agents, keep scenes whole and scene-major."""
    from cld_b200.keys import scene_buckets, scene_sizes
    sidx = torch.tensor([5] * 4 + [9] * 2 + [2] * 6 + [7] * 2 + [1] * 4)
    sizes = scene_sizes(sidx, 18)
    assert sizes == [4, 2, 6, 2, 4]
    buckets = scene_buckets(sizes)
    assert [a for a, _ in buckets] == [2, 4, 6]
    got = {a: idx.tolist() for a, idx in buckets}
    assert got[2] == [4, 5, 12, 13] and got[4] == [0, 1, 2, 3, 14, 15, 16, 17] and got[6] == list(range(6, 12))
    assert sorted(sum(got.values(), [])) == list(range(18))
    assert scene_sizes(None, 7) == [7]
    with pytest.raises(ValueError, match="not contiguous"):
        scene_sizes(torch.tensor([0, 1, 0, 1]), 4)


def test_weight_signature_sees_parent_level_loads_and_in_place_updates(models_cpu):
    """the engine's packed weights are a snapshot; the signature it is keyed on must change on a load through a
    PARENT module (Lightning's load_from_checkpoint never calls the child's load_state_dict), on optimizer steps and on
    decoder changes after VaeModel.bind."""
    dm, vae, _ = models_cpu(10)
    vae.bind(dm)

    class Parent(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dm, self.vae = dm, vae
    parent = Parent()
    s0 = dm._weights_signature()
    assert dm._weights_signature() == s0
    sd = {k: v.clone() for k, v in parent.state_dict().items()}
    parent.load_state_dict(sd)                                    # parent-level load: values equal, versions bumped
    s1 = dm._weights_signature()
    assert s1 != s0
    with torch.no_grad():
        dm.model.final_conv[1].bias.add_(1.0)                     # what an optimizer step does
    s2 = dm._weights_signature()
    assert s2 != s1
    with torch.no_grad():
        vae.lstmvae.lstm_dec.hid2act.bias.mul_(2.0)               # decoder changed after bind
    assert dm._weights_signature() != s2


def test_containers_construct_without_the_shared_library():
    """The parameter containers must not map libcld_b200.so (the bench's reference arm draws its weights from them);
    the library is loaded by engine.Engine only."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from cld_b200 import default_algo_config; "
            "from cld_b200.dm_model import DmModel; from cld_b200.vae import VaeModel; a = default_algo_config(); "
            "DmModel(a, {'image': (34, 224, 224)}, n_timesteps=10); VaeModel(a); "
            "assert not any(m.endswith('._lib') for m in sys.modules), [m for m in sys.modules if 'cld' in m]; "
            "assert 'libcld_b200' not in open('/proc/self/maps').read()") % ROOT
    subprocess.check_call([sys.executable, "-c", code])
