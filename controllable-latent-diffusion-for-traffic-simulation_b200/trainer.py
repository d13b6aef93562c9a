"""The PPO fine-tuning loop of the reference's `GuideDMLightningModule` (src/trainers/guide_dm_trainer.py:85-183) on the CUDA
path -- SURVEY.md sec. 8 f-2.  No Lightning: the class keeps the reference's method names (`training_step`, `ppo_update`,
`configure_optimizers`' Adam + warm-up/cosine schedule) and its attributes (`replay_buffer`, `steps_since_update`, ...).

Per `training_step` (guide_dm_trainer.py:85-125): sample with the current denoiser (`DmModel.forward`: x0, x1, log_prob_final),
decode + roll out, reward, push rows into the device-resident `ReplayBuffer`; every `update_interval` steps run `ppo_update`:
`ppo_epochs` x `ppo_update_times` minibatches of `ppo_mini_batch` rows, each
    log_p_new = dm.log_prob(x1, x0, cond, t = 0);  ratio = exp(log_p_new - log_p_old);  loss = -mean(min(ratio A, clip(ratio) A))
    loss.backward();  opt.step()
Two equivalent modes:
  fused=True   the whole minibatch iteration is `cld_ppo_grad` (denoiser forward, log-prob, surrogate, analytic backward into the
               flat gradient vector) + `cld_adam_step` (one launch over the flat parameter vector): no autograd graph, no Python
               per-tensor work;
  fused=False  the reference's own lines, unchanged: `DmModel.log_prob` is an autograd node whose backward is the CUDA backward,
               `torch.optim.Adam` steps the parameters.
"""
import math

import torch

from .critic import compute_reward
from .replay import ReplayBuffer


class FusedAdam:
    """torch.optim.Adam(params, lr, weight_decay) over DmModel's flat parameter vector (`cld_adam_step`)."""

    def __init__(self, dm, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.dm = dm
        self.base_lr, self.lr, self.betas, self.eps, self.weight_decay = float(lr), float(lr), betas, float(eps), float(weight_decay)
        self.flat, self.grad = dm.flatten_parameters()
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.step_count = 0

    def zero_grad(self):
        pass                                # the backward overwrites the gradient vector

    def step(self):
        self.step_count += 1
        eng = self.dm.train_engine(1)
        eng.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr, self.betas, self.eps,
                      self.weight_decay)
        self.dm.mark_parameters_changed()


class GraphedPPOStep:
    """One minibatch iteration of ppo_update -- cld_ppo_grad (forward, log-prob, surrogate, backward on two streams), cld_adam_step_dev
    and the one-launch weight re-pack -- captured ONCE as a CUDA graph and replayed: at the 128-row minibatch the ~280 launches of an
    update are bound by the host's launch rate, not by the GPU.  Inputs are copied into static buffers; the scalars that change between
    replays (reward baseline, learning rate, Adam step) live in device memory.  The first `eager_calls` calls run un-captured (they
    create the handle's scratch, tensor maps and kernel attributes); results are bit-identical to the un-captured path."""

    def __init__(self, dm, opt, rows, clip_eps=0.2, eager_calls=2):
        self.dm, self.opt, self.rows, self.clip, self.eager_left = dm, opt, int(rows), float(clip_eps), int(eager_calls)
        dev = opt.flat.device
        z = lambda *sh, dt=torch.float32: torch.zeros(*sh, device=dev, dtype=dt)      # noqa: E731
        self.x0, self.x1 = z(rows, dm.horizon, dm.latent_size), z(rows, dm.horizon, dm.latent_size)
        self.cond, self.t = z(rows, dm.cond_dim), z(rows, dt=torch.long)
        self.log_p_old, self.reward = z(rows), z(rows)
        self.baseline, self.lr, self.step = z(1), z(1, dt=torch.float64), z(1, dt=torch.long)
        self.graph, self.loss, self.log_p = None, None, None
        self._lr_host = self._base_host = None

    def buffers(self):
        """(x0, x1, log_p_old, reward, cond): sample the replay buffer straight into these (ReplayBuffer.sample(out=...))."""
        return self.x0, self.x1, self.log_p_old, self.reward, self.cond

    def _body(self, eng):
        log_p, loss = eng.ppo_grad(self.x1, self.x0, self.cond, self.t, self.log_p_old, self.reward, 0.0, self.dm._flat_views[1], self.clip,
                                   baseline_dev=self.baseline)
        o = self.opt
        eng.adam_step_dev(o.flat, o.grad, o.exp_avg, o.exp_avg_sq, self.lr, self.step, o.betas, o.eps, o.weight_decay)
        eng.load_unet(self.dm.model.state_dict())          # loaded handle: the one-launch re-pack
        return loss, log_p

    def __call__(self, baseline, t=None):
        """The static buffers hold the minibatch.  -> loss [1] (valid until the next call)."""
        dm, o = self.dm, self.opt
        if t is not None:
            self.t.copy_(t)
        if self.eager_left > 0:
            self.eager_left -= 1
            loss, _ = dm.ppo_minibatch_grad(self.x1, self.x0, self.cond, self.t, self.log_p_old, self.reward, baseline, self.clip)
            o.step()
            return loss.clone()
        if baseline != self._base_host:
            self.baseline.fill_(baseline)
            self._base_host = baseline
        if o.lr != self._lr_host:
            self.lr.fill_(o.lr)
            self._lr_host = o.lr
        if self.graph is None:
            eng = dm.train_engine(self.rows)                 # weights current, precision set: nothing left to do at replay time
            self.step.fill_(o.step_count)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss, self.log_p = self._body(eng)
        self.graph.replay()
        o.step_count += 1
        dm.mark_parameters_changed()                         # the sampling engine re-packs on its next use ...
        dm._train_sig = dm._weights_signature()              # ... the training handle was re-packed inside the graph
        return self.loss


def warmup_cosine(epoch, total_epochs):
    """lr factor of guide_dm_trainer.py:67-75."""
    warm = total_epochs / 3
    if epoch < warm:
        return float(epoch) / float(max(1, warm))
    progress = float(epoch - warm) / float(max(1, total_epochs - warm))
    return 0.5 * (1. + math.cos(math.pi * progress))


class GuideDMTrainer:
    def __init__(self, dm, vae, algo_config, *, batch_size, learning_rate=1e-4, weight_decay=0.0, epochs=30, fused=True,
                 ppo_epochs=10, clip_eps=0.2, buffer_max=None, sample_kw=None, generator=None, train_precision="tf32", cuda_graph=True):
        self.dm, self.vae, self.algo_config = dm, vae, algo_config
        self.batch_size = int(batch_size)
        self.num_samp = int(algo_config.num_samp)
        self.ppo_mini_batch = int(algo_config.ppo_mini_batch)
        self.ppo_update_times = int(algo_config.ppo_update_times)
        self.update_interval = int(algo_config.update_interval)
        self.ppo_epochs, self.clip_eps, self.epochs, self.fused = int(ppo_epochs), float(clip_eps), int(epochs), bool(fused)
        cap = buffer_max if buffer_max is not None else self.update_interval * self.batch_size * self.num_samp
        self.replay_buffer = ReplayBuffer(capacity=cap)
        self.steps_since_update = 0
        self.current_epoch = 0
        self.sample_kw = dict(sample_kw or {})
        self.generator = generator
        self.log = {}
        # "tf32": the training step's convolutions on the tensor pipe (tcgen05 kind::tf32, ~1e-3); "fp32": CUDA cores (1e-4 parity mode)
        dm.train_precision = train_precision
        for p in dm.model.parameters():
            p.requires_grad_(True)
        if self.fused:
            self.optimizer = FusedAdam(dm, lr=learning_rate, weight_decay=weight_decay)
        else:
            self.optimizer = torch.optim.Adam(dm.model.parameters(), lr=learning_rate, weight_decay=weight_decay)
        self._base_lr = float(learning_rate)
        self.cuda_graph = bool(cuda_graph) and self.fused
        self._graphed = None

    # ---- configure_optimizers' LambdaLR, stepped once per epoch (guide_dm_trainer.py:76-83)
    def on_epoch_end(self):
        self.current_epoch += 1
        lr = self._base_lr * warmup_cosine(self.current_epoch, self.epochs)
        if self.fused:
            self.optimizer.lr = lr
        else:
            for g in self.optimizer.param_groups:
                g['lr'] = lr

    @torch.no_grad()
    def training_step(self, batch, aux_info):
        """guide_dm_trainer.py:85-125; `aux_info` = what `vae.pre_vae(batch)` returns (cond_feat, curr_states)."""
        out = self.dm(batch, aux_info, self.algo_config, **self.sample_kw)
        x1, x0, log_prob_old = out['x1'], out['pred_traj'], out['log_prob_final']
        aux = out['aux_info']
        act = self.vae.lstmvae.lstm_dec(x0, aux['cond_feat'])
        traj = self.vae.convert_action_to_state_and_action(act, aux['curr_states'], descaled_output=True)
        B, N = traj.shape[0] // self.num_samp, self.num_samp
        traj4 = traj.reshape(B, N, *traj.shape[1:])
        reward = compute_reward(self.dm, traj4, batch, self.vae.scale_traj(traj4))
        if x1 is None:
            raise RuntimeError("the sampler did not visit step 1 (stride > 1): the PPO update needs x1 (dm_model.py:126-127)")
        self.replay_buffer.add(x0, x1, log_prob_old, reward, aux['cond_feat'])
        self.steps_since_update += 1
        self.log['train/reward'] = float(reward.mean().item())
        if self.steps_since_update >= self.update_interval:
            with torch.enable_grad():
                self.log['train/ppo_loss'] = float(self.ppo_update().item())
            self.steps_since_update = 0
        return {'traj': traj4[..., :2], 'reward': reward}

    def ppo_update(self):
        """guide_dm_trainer.py:127-183."""
        losses = []
        for _ in range(self.ppo_epochs):
            for _ in range(self.ppo_update_times):
                baseline = self.replay_buffer.get_baseline()
                if self.cuda_graph:
                    if self._graphed is None:
                        self._graphed = GraphedPPOStep(self.dm, self.optimizer, self.ppo_mini_batch, self.clip_eps)
                    self.replay_buffer.sample(self.ppo_mini_batch, self.generator, out=self._graphed.buffers())
                    losses.append(self._graphed(baseline).reshape(()).clone())      # t stays 0 (guide_dm_trainer.py:160)
                    continue
                x0, x1, log_p_old, reward, cond = self.replay_buffer.sample(self.ppo_mini_batch, self.generator)
                t = torch.zeros(x0.shape[0], device=x0.device, dtype=torch.long)
                losses.append(self.ppo_minibatch(x0, x1, log_p_old, reward, cond, t, baseline))
        return torch.stack(losses).mean()

    def ppo_minibatch(self, x0, x1, log_p_old, reward, cond, t, baseline):
        if self.fused:
            loss, _ = self.dm.ppo_minibatch_grad(x1, x0, cond, t, log_p_old, reward, baseline, self.clip_eps)
            self.optimizer.step()
            return loss.reshape(())
        advantage = reward - baseline
        log_p_new = self.dm.log_prob(x1, x0, {'cond_feat': cond}, t=t)
        ratios = torch.exp(log_p_new - log_p_old)
        surr1 = ratios * advantage
        surr2 = torch.clamp(ratios, 1 - self.clip_eps, 1 + self.clip_eps) * advantage
        loss = -torch.min(surr1, surr2).mean()
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.detach()
