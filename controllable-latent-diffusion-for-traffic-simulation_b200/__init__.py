"""cld_b200: B200-native guided latent-diffusion sampling path of CLD (drop-in behind DmModel).

Importing the configuration / synthetic-data helpers needs no GPU; anything that computes goes through
libcld_b200.so and fails loudly when it is missing or the device is not sm_100.
"""
from .config import ConfigBase, default_algo_config, dict_to_config  # noqa: F401
from .synthetic import make_context_batch, make_scenes  # noqa: F401


def __getattr__(name):
    # lazy: these import the shared library
    if name in ("DmModel", "TemporalMapUnetParams", "cosine_beta_schedule"):
        from . import dm_model
        return getattr(dm_model, name)
    if name in ("VaeModel", "LSTMVAE"):
        from . import vae
        return getattr(vae, name)
    if name in ("Engine", "default_guidance", "DECODER_KEYS"):
        from . import engine
        return getattr(engine, name)
    if name == "ContextEncoder":
        from . import context
        return context.ContextEncoder
    if name in ("ReplayBuffer", "ppo_surrogate"):
        from . import replay
        return getattr(replay, name)
    if name in ("GuideDMTrainer", "FusedAdam", "GraphedPPOStep", "warmup_cosine"):
        from . import trainer
        return getattr(trainer, name)
    if name in ("TargetPosAtTime", "GlobalTargetPosAtTime", "GlobalTargetPos"):
        from . import waypoints
        return getattr(waypoints, name)
    if name in ("SyntheticEnv", "closed_loop_rollout"):
        from . import rollout
        return getattr(rollout, name)
    if name in ("GuidedDiffusionPolicy", "choose_action_from_guidance"):
        from . import policy
        return getattr(policy, name)
    if name == "HostStager":
        from . import staging
        return staging.HostStager
    if name in ("failure_rate_compute", "compute_reward", "indicators"):
        from . import critic
        return getattr(critic, name)
    if name in ("RealismAccumulator", "wasserstein_1d", "realism_samples", "validation_step"):
        from . import metrics
        return getattr(metrics, name)
    raise AttributeError(name)
