"""Synthetic nuScenes-shaped scenes for tests and bench.

Shapes follow the reference's trajdata batch as parsed by `parse_node_centric`
(src/tbsim/utils/trajdata_utils.py:346-475): agent-centric frames, 224x224 raster at 0.5 m/px with
`raster_from_agent = [[2,0,56],[0,2,112],[0,0,1]]` (trajdata_utils.py:385-389, config.yaml:73-86).
All tensors are CPU fp32/bool; callers move them to the device.
"""
import math

import torch


def pack_drivable_map(dmap):
    """[B,H,W] bool / uint8 -> [B,H,(W+7)//8] uint8, pixel x = bit (x & 7) of byte x >> 3 (numpy.packbits bitorder="little"): the
    format `Engine.make_scene` accepts as data_batch["drivable_map_bits"] -- 1/8 of the bytes to ship per agent."""
    d = dmap.to(torch.uint8)
    B, H, W = d.shape
    if W % 8:
        d = torch.nn.functional.pad(d, (0, 8 - W % 8))
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=d.device)
    return (d.reshape(B, H, -1, 8) * w).sum(-1).to(torch.uint8).contiguous()


def make_scenes(num_scenes, agents_per_scene, horizon=52, seed=123, cond_dim=256, dense=False):
    """Returns (aux_info, batch).  B = num_scenes * agents_per_scene agent rows, scene-major.

    dense=True packs the agents of a scene within ~12 m so that agent-agent collisions and
    off-road events actually occur (used by the guidance tests)."""
    g = torch.Generator().manual_seed(seed)
    S, A, T = num_scenes, agents_per_scene, horizon
    B = S * A
    cond = torch.randn(B, cond_dim, generator=g)
    v = torch.rand(B, generator=g) * 15.0
    stationary = torch.rand(B, generator=g) < 0.2
    v = torch.where(stationary, torch.rand(B, generator=g) * 0.4, v)
    curr = torch.stack([torch.zeros(B), torch.zeros(B), v, torch.zeros(B)], dim=1)
    extent = torch.stack([4.0 + 1.5 * torch.rand(B, generator=g),
                          1.8 + 0.4 * torch.rand(B, generator=g),
                          torch.full((B,), 1.6)], dim=1)
    spread = 12.0 if dense else 50.0
    pos = (torch.rand(B, 2, generator=g) * 2 - 1) * spread
    if dense:
        yaw = (torch.rand(B, generator=g) * 2 - 1) * 0.3 + \
            math.pi * (torch.rand(B, generator=g) < 0.5).float()
    else:
        yaw = (torch.rand(B, generator=g) * 2 - 1) * math.pi
    c, s = torch.cos(yaw), torch.sin(yaw)
    wfa = torch.zeros(B, 3, 3)
    wfa[:, 0, 0], wfa[:, 0, 1], wfa[:, 0, 2] = c, -s, pos[:, 0]
    wfa[:, 1, 0], wfa[:, 1, 1], wfa[:, 1, 2] = s, c, pos[:, 1]
    wfa[:, 2, 2] = 1.0
    rfa = torch.tensor([[2., 0., 56.], [0., 2., 112.], [0., 0., 1.]]).repeat(B, 1, 1)
    # drivable map: a road band around the agent's heading axis + random rectangles
    dmap = torch.zeros(B, 224, 224, dtype=torch.bool)
    half = torch.randint(6, 22, (B,), generator=g)          # band half-width in px (3..11 m)
    off = torch.randint(-6, 7, (B,), generator=g)
    rows = torch.arange(224)[None, :]
    band = (rows >= (112 + off - half)[:, None]) & (rows <= (112 + off + half)[:, None])
    dmap |= band[:, :, None]
    nrect = 3
    r0 = torch.randint(0, 200, (B, nrect), generator=g)
    c0 = torch.randint(0, 200, (B, nrect), generator=g)
    rh = torch.randint(8, 60, (B, nrect), generator=g)
    cw = torch.randint(8, 60, (B, nrect), generator=g)
    cols = torch.arange(224)[None, :]
    for k in range(nrect):
        rm = (rows >= r0[:, k:k + 1]) & (rows < (r0[:, k:k + 1] + rh[:, k:k + 1]))
        cm = (cols >= c0[:, k:k + 1]) & (cols < (c0[:, k:k + 1] + cw[:, k:k + 1]))
        dmap |= rm[:, :, None] & cm[:, None, :]
    # other agents' futures in each ego frame: straight-line constant-velocity motion
    tt = (torch.arange(1, T + 1).float() * 0.1)
    fut_local = torch.stack([v[:, None] * tt[None, :], torch.zeros(B, T)], dim=-1)          # [B,T,2]
    fut_world = torch.einsum('bij,btj->bti', wfa[:, :2, :2], fut_local) + wfa[:, None, :2, 2]
    fw = fut_world.view(S, A, T, 2)
    So = max(A - 1, 1)
    others = torch.zeros(S, A, So, T, 2)
    for a in range(A):
        idx = [j for j in range(A) if j != a] or [a]
        others[:, a] = fw[:, idx]
    others = others.view(B, So, T, 2)
    Rinv = wfa[:, :2, :2].transpose(1, 2)
    others = torch.einsum('bij,bstj->bsti', Rinv, others - wfa[:, None, None, :2, 2])
    avail = torch.rand(B, So, T, generator=g) < 0.7
    if A == 1:
        avail[:] = False
    target = torch.stack([v * 5.2, (torch.rand(B, generator=g) * 2 - 1) * 3.0], dim=1)
    aux = {'cond_feat': cond, 'curr_states': curr}
    batch = {
        'history_positions': torch.zeros(B, 31, 2),
        'curr_speed': v.clone(),
        'extent': extent,
        'world_from_agent': wfa,
        'raster_from_agent': rfa,
        'scene_index': torch.arange(S).repeat_interleave(A),
        'drivable_map': dmap,
        'all_other_agents_future_positions': others.contiguous(),
        'all_other_agents_future_availability': avail,
        'target_pos': target,
    }
    return aux, batch


def make_context_batch(num_agents, seed=321, hist=31, size=224):
    """Synthetic inputs of the context encoder (models/context_utils.py:40-61), shaped like trajdata's agent-centric
    raster (src/tbsim/utils/trajdata_utils.py:123-156): channels 0..30 are the history frames (ego pixel +1, other agents
    -1, everything else 0), channels 31..33 the map layers in [0,1].  Every image value is a multiple of 0.5, so a
    golden file can keep the raster as int8.  Returns the `data_batch` entries ContextEncoder.forward reads."""
    g = torch.Generator().manual_seed(seed)
    B = num_agents
    img = torch.zeros(B, hist + 3, size, size)
    v = torch.rand(B, generator=g) * 15.0
    t_back = torch.arange(hist - 1, -1, -1).float() * 0.1                    # seconds before now, oldest first
    hist_pos = torch.stack([-v[:, None] * t_back[None, :], 0.2 * torch.randn(B, hist, generator=g)], dim=-1)
    hist_pos[:, -1] = 0.0
    hist_yaw = 0.05 * torch.randn(B, hist, 1, generator=g)
    bi = torch.arange(B)
    for t in range(hist):
        px = (hist_pos[:, t, 0] * 2 + 56).round().long().clamp(0, size - 1)
        py = (hist_pos[:, t, 1] * 2 + 112).round().long().clamp(0, size - 1)
        n_other = 6
        ox = torch.randint(0, size, (B, n_other), generator=g)
        oy = torch.randint(0, size, (B, n_other), generator=g)
        for k in range(n_other):
            img[bi, t, oy[:, k], ox[:, k]] = -1.0
        img[bi, t, py, px] = 1.0
    rows = torch.arange(size)[None, :, None]
    cols = torch.arange(size)[None, None, :]
    for c in range(3):
        half = torch.randint(6, 40, (B,), generator=g)[:, None, None]
        off = torch.randint(-20, 21, (B,), generator=g)[:, None, None]
        band = ((rows >= 112 + off - half) & (rows <= 112 + off + half)).expand(B, size, size)
        r0 = torch.randint(0, 180, (B,), generator=g)[:, None, None]
        c0 = torch.randint(0, 180, (B,), generator=g)[:, None, None]
        rect = (rows >= r0) & (rows < r0 + 44) & (cols >= c0) & (cols < c0 + 44)
        img[:, hist + c] = band.float() * (1.0 if c != 1 else 0.5) + rect.float() * 0.5
    img[:, hist:].clamp_(0.0, 1.0)
    return {'image': img, 'history_positions': hist_pos, 'history_yaws': hist_yaw, 'curr_speed': v}


def make_history_batch(num_agents, num_neighbors=5, seed=77, hist=31, size=224):
    """Inputs of the history rasteriser (src/tbsim/utils/trajdata_utils.py:123-156) as `parse_node_centric` assembles them
    (:395-420): map layers [B,3,H,W] with values in {0, 0.5, 1}, history positions of the ego (index 0) and its neighbours
    in the ego frame [B,1+Nn,T,2], their availability mask and `raster_from_agent`.  Some positions fall outside the raster,
    some are unavailable, several agents share a pixel."""
    g = torch.Generator().manual_seed(seed)
    B, A = num_agents, 1 + num_neighbors
    rows = torch.arange(size)[None, :, None]
    cols = torch.arange(size)[None, None, :]
    maps = torch.zeros(B, 3, size, size)
    for c in range(3):
        half = torch.randint(6, 40, (B,), generator=g)[:, None, None]
        off = torch.randint(-20, 21, (B,), generator=g)[:, None, None]
        band = ((rows >= 112 + off - half) & (rows <= 112 + off + half)).expand(B, size, size)
        c0 = torch.randint(0, 180, (B,), generator=g)[:, None, None]
        rect = (cols >= c0) & (cols < c0 + 44)
        maps[:, c] = (band.float() * (1.0 if c != 1 else 0.5) + (band & rect).float() * 0.5).clamp(0, 1)
    v = torch.rand(B, A, generator=g) * 12.0
    t_back = torch.arange(hist - 1, -1, -1).float() * 0.1
    start = (torch.rand(B, A, 2, generator=g) * 2 - 1) * torch.tensor([70.0, 60.0])       # some start outside the 112 m raster
    start[:, 0] = 0.0
    pos = start[:, :, None, :] + torch.stack([-v[:, :, None] * t_back[None, None, :], torch.zeros(B, A, hist)], dim=-1)
    pos[:, 1] = pos[:, 2]                                                               # two neighbours on the same pixels
    mask = torch.rand(B, A, hist, generator=g) < 0.8
    mask[:, 0] = True
    rfa = torch.tensor([[2., 0., 56.], [0., 2., 112.], [0., 0., 1.]]).repeat(B, 1, 1)
    return {'maps': maps, 'agent_hist_pos': pos, 'agent_hist_mask': mask, 'raster_from_agent': rfa}
