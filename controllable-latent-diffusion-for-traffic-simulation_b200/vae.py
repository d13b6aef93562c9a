"""Drop-in for the parts of the reference's VaeModel that sit on the sampling path
(models/vae/vae_model.py, models/vae/lstm_vae.py): the LSTM decoder, the action -> state rollout and
the (de)scaling helpers.  `lstmvae.*` parameter names match the reference so `vae.lstmvae.*`
checkpoint keys load unchanged.  `context_encoder` (models/vae/vae_model.py:48) is created when
`modality_shapes` is given, as in the reference; `pre_vae`'s context part is `self.context_encoder(batch)`.
"""
import numpy as np
import torch
import torch.nn as nn

from .keys import DECODER_KEYS


class _LstmBlock(nn.Module):
    """Encoder / Decoder containers (lstm_vae.py:6-52): construction order lstm, cond2hidden[, hid2act]."""

    def __init__(self, input_size, hidden_size, num_layers, output_size=None, cond_dim=256, dropout_rate=0.2):
        super().__init__()
        self.hidden_size, self.num_layers = hidden_size, num_layers
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers, batch_first=True, bidirectional=False,
                            dropout=dropout_rate if num_layers > 1 else 0.0)
        self.cond2hidden = nn.Linear(cond_dim, hidden_size)
        if output_size is not None:
            self.hid2act = nn.Linear(hidden_size, output_size)
        self._owner = None

    def forward(self, x, context):
        """Decoder.forward(z, cond) -> scaled actions [R,T,2] on the B200 kernel."""
        if not hasattr(self, "hid2act"):
            raise RuntimeError("the VAE encoder is training-side and not part of the sampling path")
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise RuntimeError("decoder container is not attached to a VaeModel")
        act, _ = owner._decode_rollout(x, context, None)
        return act


class LSTMVAE(nn.Module):
    def __init__(self, input_size, hidden_size, latent_size, output_size, dropout_rate=0.2):
        super().__init__()
        self.input_size, self.hidden_size, self.latent_size, self.num_layers = input_size, hidden_size, latent_size, 2
        self.lstm_enc = _LstmBlock(input_size, hidden_size, 2, None, dropout_rate=dropout_rate)
        self.lstm_dec = _LstmBlock(latent_size, hidden_size, 2, output_size, dropout_rate=dropout_rate)
        self.mu = nn.Linear(hidden_size, latent_size)
        self.logvar = nn.Linear(hidden_size, latent_size)


class VaeModel(nn.Module):
    """`VaeModel(algo_config, train_config, modality_shapes)` -- sampling-path subset."""

    def __init__(self, algo_config, train_config=None, modality_shapes=None, *, dm=None):
        super().__init__()
        import weakref
        self.algo_config = algo_config
        vae_config = algo_config.vae
        self.lstmvae = LSTMVAE(input_size=6, hidden_size=vae_config.hidden_size, latent_size=vae_config.latent_size,
                               output_size=2)
        self.lstmvae.lstm_dec._owner = weakref.ref(self)
        self.default_chosen_inds = [0, 1, 2, 3, 4, 5]
        norm = algo_config.nusc_norm_info.diffuser
        self.add_coeffs = np.array(norm[0]).astype('float32')
        self.div_coeffs = np.array(norm[1]).astype('float32')
        self.horizon = algo_config.horizon
        self.dt = 0.1
        self._dm = None
        if dm is not None:
            self.bind(dm)
        if modality_shapes is not None:
            # reference order (vae_model.py:28-52): lstmvae first, the context encoder last
            from .context import ContextEncoder
            self.context_encoder = ContextEncoder(4, algo_config, modality_shapes, None)

    def bind(self, dm):
        """Share the DmModel's engine (one handle per device) and hand it the decoder weights."""
        object.__setattr__(self, "_dm", dm)
        dm.attach_decoder(self.decoder_state_dict(), module=self.lstmvae.lstm_dec)
        return self

    def decoder_state_dict(self):
        sd = self.lstmvae.lstm_dec.state_dict()
        return {k: sd[k] for k in DECODER_KEYS}

    def _engine(self, rows):
        if self._dm is None:
            raise RuntimeError("VaeModel.bind(dm) must be called before decoding (the decoder runs on the DmModel's engine)")
        return self._dm.engine(rows)

    def _decode_rollout(self, z, cond, curr):
        R = z.shape[0]
        if curr is None:
            curr = torch.zeros(R, 4, device=z.device)
        return self._engine(R).decode_rollout(z, cond, curr)

    @torch.no_grad()
    def decode_to_trajectory(self, z, cond, curr_states):
        """lstm_dec + convert_action_to_state_and_action(descaled_output=True) in ONE kernel:
        the three lines after `self.dm(...)` in guide_dm_trainer.py:88-90."""
        act, traj = self._decode_rollout(z, cond, curr_states)
        return traj, act

    @torch.no_grad()
    def convert_action_to_state_and_action(self, x_out, curr_states, scaled_input=True, descaled_output=False):
        """vae_model.py:100-129 (unicycle on the B200 kernel)."""
        dim = x_out.dim()
        if dim == 4:
            B, N, T, _ = x_out.shape
            x_out = x_out.reshape(B * N, T, -1)
        if scaled_input:
            x_out = self.descale_traj(x_out, [4, 5])
        state = self._engine(x_out.shape[0]).unicycle(curr_states, x_out)
        x_all = torch.cat([state, x_out], dim=-1)
        if scaled_input and not descaled_output:
            x_all = self.scale_traj(x_all, [0, 1, 2, 3, 4, 5])
        if dim == 4:
            x_all = x_all.reshape(B, N, T, -1)
        return x_all

    def scale_traj(self, traj, chosen_inds=[]):
        inds = chosen_inds if len(chosen_inds) else self.default_chosen_inds
        squeeze = traj.dim() == 2
        if squeeze:
            traj = traj.unsqueeze(1)
        mean = torch.tensor(self.add_coeffs[inds][None, None], device=traj.device)
        std = torch.tensor(self.div_coeffs[inds][None, None], device=traj.device)
        out = (traj - mean) / std
        return out.squeeze(1) if squeeze else out

    def descale_traj(self, traj, chosen_inds=[]):
        inds = chosen_inds if len(chosen_inds) else self.default_chosen_inds
        mean = torch.tensor(self.add_coeffs[inds][None, None], device=traj.device)
        std = torch.tensor(self.div_coeffs[inds][None, None], device=traj.device)
        return traj * std + mean
