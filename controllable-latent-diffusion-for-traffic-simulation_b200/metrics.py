"""Evaluation statistics of the guided sampler's output, device-resident.

Mirrors the test-time half of the reference's trainer:
  * `realism_samples`      GuideDMLightningModule.test_step                 src/trainers/guide_dm_trainer.py:205-247
  * `wasserstein_1d`       scipy.stats.wasserstein_distance (unweighted)    called at src/trainers/guide_dm_trainer.py:277-279
  * `RealismAccumulator`   test_outputs + on_test_epoch_end                 src/trainers/guide_dm_trainer.py:236-296

The reference copies every batch's accelerations and jerks to the host as numpy arrays and calls scipy on the concatenation.  Here the
samples stay on the device they were produced on (the sampler's), the distance is one sort + two binary searches per pair of
distributions, and only three scalars cross to the host.  The arithmetic follows scipy's `_cdf_distance(p=1)`: with the pooled, sorted
values x_0 <= ... <= x_n and the empirical CDFs U, V evaluated right-continuously at x_0 .. x_{n-1},
    W1 = sum_i |U(x_i) - V(x_i)| * (x_{i+1} - x_i),
accumulated in float64 as scipy does.  (The torch reference of this repository's CUDA kernels is `oracle/`; this module is host
logic on torch tensors and needs no kernel of its own: it is bandwidth-trivial next to the sampler.)
"""
import torch


def wasserstein_1d(u_values, v_values):
    """First Wasserstein distance between two 1-D empirical distributions (equal weights); a 0-d float64 tensor on the inputs' device."""
    u = torch.as_tensor(u_values).reshape(-1).to(torch.float64)
    v = torch.as_tensor(v_values).reshape(-1).to(u.device, torch.float64)
    if u.numel() == 0 or v.numel() == 0:
        raise ValueError("wasserstein_1d needs non-empty samples")
    if not (torch.isfinite(u).all() and torch.isfinite(v).all()):
        raise ValueError("wasserstein_1d: samples must be finite")
    u_sorted, v_sorted = torch.sort(u).values, torch.sort(v).values
    all_values = torch.sort(torch.cat([u, v])).values
    deltas = all_values[1:] - all_values[:-1]
    # right-continuous empirical CDFs at every pooled value but the last
    u_cdf = torch.searchsorted(u_sorted, all_values[:-1], right=True).to(torch.float64) / u.numel()
    v_cdf = torch.searchsorted(v_sorted, all_values[:-1], right=True).to(torch.float64) / v.numel()
    return ((u_cdf - v_cdf).abs() * deltas).sum()


def realism_samples(pred_state_and_action, gt_state_and_action, dt):
    """The six sample sets of test_step (guide_dm_trainer.py:219-247) from SCALED [R, T, 6] trajectories (x, y, v, yaw, acc, yaw rate):
    longitudinal acceleration = channel 4, lateral acceleration = v * yaw rate (channels 2 * 5), jerk = forward difference of the
    longitudinal acceleration over dt.  Returned flat, on the inputs' device."""
    p, g = torch.as_tensor(pred_state_and_action), torch.as_tensor(gt_state_and_action)
    if p.shape[-1] < 6 or g.shape[-1] < 6 or p.dim() != 3 or g.dim() != 3:
        raise ValueError("expected [R, T, 6] state-and-action tensors")
    long_p, long_g = p[..., 4], g[..., 4]
    return {
        "long_acc_gt": long_g.reshape(-1), "long_acc_pred": long_p.reshape(-1),
        "lat_acc_gt": (g[..., 2] * g[..., 5]).reshape(-1), "lat_acc_pred": (p[..., 2] * p[..., 5]).reshape(-1),
        "jerk_gt": ((long_g[:, 1:] - long_g[:, :-1]) / dt).reshape(-1), "jerk_pred": ((long_p[:, 1:] - long_p[:, :-1]) / dt).reshape(-1),
    }


class RealismAccumulator:
    """Collects the per-batch sample sets and failure statistics of a test epoch and reduces them as on_test_epoch_end does
    (guide_dm_trainer.py:252-296).  NOTE the reference re-creates `test_outputs` / `all_failure_stats` inside every test_step
    (:211, :236), so its epoch-end numbers cover the LAST batch only; `last_batch_only=True` reproduces that, the default keeps
    every batch (the evident intent of the list-and-concatenate code)."""

    def __init__(self, dt, last_batch_only=False):
        self.dt = float(dt)
        self.last_batch_only = bool(last_batch_only)
        self._samples, self._stats = [], []

    def add_batch(self, pred_state_and_action_scaled, gt_state_and_action_scaled, failure_stats=None):
        if self.last_batch_only:
            self._samples, self._stats = [], []
        self._samples.append(realism_samples(pred_state_and_action_scaled, gt_state_and_action_scaled, self.dt))
        if failure_stats is not None:
            self._stats.append(failure_stats)

    def compute(self):
        if not self._samples:
            raise RuntimeError("no batch added")
        cat = {k: torch.cat([s[k] for s in self._samples]) for k in self._samples[0]}
        wd_long = wasserstein_1d(cat["long_acc_gt"], cat["long_acc_pred"])
        wd_lat = wasserstein_1d(cat["lat_acc_gt"], cat["lat_acc_pred"])
        wd_jerk = wasserstein_1d(cat["jerk_gt"], cat["jerk_pred"])
        out = {"wd_long": wd_long.item(), "wd_lat": wd_lat.item(), "wd_jerk": wd_jerk.item()}
        out["realism_deviation"] = (out["wd_long"] + out["wd_lat"] + out["wd_jerk"]) / 3.0
        if self._stats:
            for k in ("offroad_failure_rate", "collision_failure_rate", "overall_failure_rate"):
                vals = [float(s[k]) for s in self._stats if k in s]
                if vals:
                    out["avg_" + k] = sum(vals) / len(vals)
        return out


@torch.no_grad()
def test_step(dm, vae, batch, aux_info, state_and_action, algo_config, accumulator, **sample_kw):
    """GuideDMLightningModule.test_step (guide_dm_trainer.py:205-247) on the B200 path: sample, decode + roll out, failure rates,
    realism sample sets into `accumulator`.  `state_and_action` is the SCALED ground-truth [B, T, 6] tensor `pre_vae` returns in the
    reference; `sample_kw` goes to DmModel.forward (sampler, guidance, seed ...).  Returns the failure statistics of the batch."""
    from .critic import failure_rate_compute
    out = dm(batch, aux_info, algo_config, **sample_kw)
    x0, aux = out["pred_traj"], out["aux_info"]
    action = vae.lstmvae.lstm_dec(x0, aux["cond_feat"])
    recon_descaled = vae.convert_action_to_state_and_action(action, aux["curr_states"], descaled_output=True)
    stats = failure_rate_compute(dm, recon_descaled, batch)
    recon_scaled = vae.scale_traj(recon_descaled)
    accumulator.add_batch(recon_scaled, state_and_action.to(recon_scaled.device), stats)
    return stats


test_step.__test__ = False        # not a pytest test


@torch.no_grad()
def validation_step(dm, vae, batch, aux_info, algo_config, **sample_kw):
    """GuideDMLightningModule.validation_step (guide_dm_trainer.py:185-202) on the B200 path: sample, decode + roll out, mean reward
    (`val/reward`).  Returns a 0-d tensor on the sampler's device."""
    from .critic import compute_reward
    out = dm(batch, aux_info, algo_config, **sample_kw)
    x0, aux = out["pred_traj"], out["aux_info"]
    action = vae.lstmvae.lstm_dec(x0, aux["cond_feat"])
    recon_descaled = vae.convert_action_to_state_and_action(action, aux["curr_states"], descaled_output=True)
    N = int(algo_config.num_samp)
    B = recon_descaled.shape[0] // N
    state = recon_descaled.reshape(B, N, *recon_descaled.shape[1:])
    return compute_reward(dm, state, batch).mean()
