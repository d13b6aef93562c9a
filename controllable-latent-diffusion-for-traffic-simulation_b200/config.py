"""Attribute + item access config objects, equivalent to the reference's ConfigBase
(configs/custom_config.py:1-41), and the default algo config of config.yaml."""


class ConfigBase:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, ConfigBase(**v) if isinstance(v, dict) else v)

    def get(self, key, default=None):
        return getattr(self, key, default)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigBase) else v) for k, v in self.__dict__.items()}

    def items(self):
        return self.to_dict().items()

    def __contains__(self, key):
        return key in self.__dict__

    def __getitem__(self, key):
        if key in self.__dict__:
            return self.__dict__[key]
        raise KeyError("Key '%s' not found in ConfigBase." % key)

    def __setitem__(self, key, value):
        self.__dict__[key] = ConfigBase(**value) if isinstance(value, dict) else value


def dict_to_config(d):
    return ConfigBase(**d)


def default_algo_config(**over):
    """The `algo:` section keys of the reference's config.yaml that the sampling path reads."""
    d = dict(
        name="dm_vae", horizon=52, step_time=0.1, base_dim=32, dim_mults=[2, 4, 8], cond_feat_dim=256,
        curr_state_feat_dim=64, map_feature_dim=256, map_encoder_model_arch="resnet18", n_diffusion_steps=100,
        vae=dict(hidden_size=64, latent_size=4),
        dynamics=dict(type="Unicycle", max_steer=0.5, max_yawvel=6.283185307179586, acce_bound=[-10, 8],
                      ddh_bound=[-6.283185307179586, 6.283185307179586], max_speed=40.0),
        nusc_norm_info=dict(diffuser=[[13.162, -0.13891, 5.0223, -0.0046415, -0.0080072, -0.0013546],
                                      [13.0717, 2.2462, 3.6187, 0.2210, 2.5770, 0.0840]]),
        num_samp=1,
        # PPO fine-tuning (config.yaml:167-172)
        ppo_mini_batch=128, ppo_update_times=300, update_interval=10,
    )
    d.update(over)
    return ConfigBase(**d)
