"""Drop-ins for models/rl/criticmodel.py on the sampling path: failure_rate_compute, compute_reward."""
import torch


@torch.no_grad()
def indicators(dm, state_action, batch, num_samp=1):
    """Per-row indicators: offroad [R,T] bool, collision counts [R], reward [R].  Ragged batches (scenes of different sizes) are
    evaluated as one uniform sub-batch per scene size (`keys.scene_buckets`)."""
    from .keys import scene_buckets, scene_sizes
    R = state_action.shape[0]
    x = state_action
    if x.shape[-1] < 6:
        x = torch.cat([x, x.new_zeros(*x.shape[:-1], 6 - x.shape[-1])], dim=-1)
    B = R // num_samp
    sizes = scene_sizes(batch.get('scene_index'), B)
    if len(set(sizes)) == 1:
        A = sizes[0]
        return dm.engine(R).indicators(x, dm.engine(R).make_scene(batch, B // A, A, num_samp))
    outs = None
    for A, idx in scene_buckets(sizes):
        idx = idx.to(x.device)
        rows = (idx[:, None] * num_samp + torch.arange(num_samp, device=x.device)[None, :]).reshape(-1)
        nb = idx.numel()
        sub = {k: (v.index_select(0, idx.to(v.device)) if (torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B) else v) for k, v in batch.items()}
        sub['scene_index'] = torch.arange(nb // A).repeat_interleave(A)
        eng = dm.engine(nb * num_samp)
        o = eng.indicators(x.index_select(0, rows), eng.make_scene(sub, nb // A, A, num_samp))
        if outs is None:
            outs = [v.new_empty((R,) + tuple(v.shape[1:])) for v in o]
        for dst, v in zip(outs, o):
            dst.index_copy_(0, rows, v)
    return tuple(outs)


@torch.no_grad()
def failure_rate_compute(dm, state_action, batch):
    """criticmodel.py:114-145."""
    off, coll, _ = indicators(dm, state_action, batch, 1)
    no_off = (~off).all(dim=-1).float().mean().item()
    no_col = (coll <= 0).float().mean().item()
    o, c = 1.0 - no_off, 1.0 - no_col
    return {'offroad_failure_rate': o, 'collision_failure_rate': c, 'overall_failure_rate': (o + c) / 2.0}


@torch.no_grad()
def compute_reward(dm, state_act, batch, state_act_scaled=None):
    """criticmodel.py:7-40 with its evident [B,N,T,6] intent; returns [B*N]."""
    B, N, T, _ = state_act.shape
    _, _, rew = indicators(dm, state_act.reshape(B * N, T, -1), batch, N)
    return rew
