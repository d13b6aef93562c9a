"""Closed-loop rollout driver: the loop of the reference's `rollout.py:95-100`

    done = env.is_done()
    while not done:
        obs = env.get_observation()
        action = policy.get_action(obs)
        env.step(action)
        done = env.is_done()

on top of `GuidedDiffusionPolicy.get_action` (the B200 sampler).  The reference's environment is trajdata's simulation scene
(`EnvUnifiedBuilder`, nuScenes maps) -- absent here and out of scope; `SyntheticEnv` is the stand-in the survey asks for ("needs
trajdata / nuScenes -> synthetic env first"): S scenes of A agents driving on a straight road in a WORLD frame, observations
re-centred on every agent at every replanning step in the layout `parse_node_centric` produces (`trajdata_utils.py:346-475`:
agent-centric history, world_from_agent / agent_from_world, raster_from_agent, the drivable raster of the agent's own frame, the
other agents' futures in the ego frame).  Everything is device-resident torch plumbing; all arithmetic of the policy happens in
libcld_b200.
"""
import math

import torch


class SyntheticEnv:
    def __init__(self, num_scenes, agents_per_scene, *, horizon=52, dt=0.1, history=31, road_half_width=7.0, num_steps=100,
                 n_step_action=10, seed=0, device="cuda"):
        self.S, self.A, self.T, self.dt, self.H = int(num_scenes), int(agents_per_scene), int(horizon), float(dt), int(history)
        self.half, self.num_steps, self.n_step_action = float(road_half_width), int(num_steps), int(n_step_action)
        self.device, self.seed = torch.device(device), int(seed)
        self.reset()

    # ------------------------------------------------------------------ state
    def reset(self):
        g = torch.Generator().manual_seed(self.seed)
        B = self.S * self.A
        lane = (torch.rand(B, generator=g) * 2 - 1) * (self.half - 1.5)
        x = torch.arange(self.A).float().repeat(self.S) * 9.0 + torch.rand(B, generator=g) * 3.0       # staggered along the road
        head = math.pi * (torch.rand(B, generator=g) < 0.3).float() + (torch.rand(B, generator=g) * 2 - 1) * 0.05
        speed = torch.rand(B, generator=g) * 9.0 + 1.0
        speed = torch.where(torch.rand(B, generator=g) < 0.15, torch.zeros(B), speed)                 # some parked agents
        d = self.device
        self.pos = torch.stack([x, lane], 1).to(d)                  # world x, y
        self.yaw, self.speed = head.to(d), speed.to(d)
        self.extent = torch.stack([4.0 + 1.5 * torch.rand(B, generator=g), 1.8 + 0.4 * torch.rand(B, generator=g),
                                   torch.full((B,), 1.6)], 1).to(d)
        # world history: a constant-velocity past
        k = torch.arange(self.H - 1, -1, -1, device=d).float() * self.dt                               # oldest first
        dirv = torch.stack([torch.cos(self.yaw), torch.sin(self.yaw)], 1)
        self.hist_pos = self.pos[:, None, :] - dirv[:, None, :] * (self.speed[:, None, None] * k[None, :, None])
        self.hist_yaw = self.yaw[:, None].repeat(1, self.H)
        self.plan_world = self.pos[:, None, :] + dirv[:, None, :] * (self.speed[:, None, None] *
                                                                    (torch.arange(1, self.T + 1, device=d).float() * self.dt)[None, :, None])
        self.t = 0
        self.offroad_steps = torch.zeros(B, device=d)
        self.collision_steps = torch.zeros(B, device=d)

    def is_done(self):
        return self.t >= self.num_steps

    # ------------------------------------------------------------------ frames
    def _world_from_agent(self):
        c, s = torch.cos(self.yaw), torch.sin(self.yaw)
        B = c.shape[0]
        m = torch.zeros(B, 3, 3, device=self.device)
        m[:, 0, 0], m[:, 0, 1], m[:, 0, 2] = c, -s, self.pos[:, 0]
        m[:, 1, 0], m[:, 1, 1], m[:, 1, 2] = s, c, self.pos[:, 1]
        m[:, 2, 2] = 1.0
        return m

    def _to_agent(self, pts_world, wfa):
        """pts_world [B, ..., 2] -> the frame of agent b."""
        R, p = wfa[:, :2, :2], wfa[:, :2, 2]
        shp = (-1,) + (1,) * (pts_world.dim() - 2) + (2,)
        return torch.einsum('bji,b...j->b...i', R, pts_world - p.reshape(shp))

    def drivable_world(self, xy):
        return xy[..., 1].abs() <= self.half

    # ------------------------------------------------------------------ observation (parse_node_centric layout)
    def get_observation(self):
        d, B, S, A, T = self.device, self.S * self.A, self.S, self.A, self.T
        wfa = self._world_from_agent()
        rfa = torch.tensor([[2., 0., 56.], [0., 2., 112.], [0., 0., 1.]], device=d).repeat(B, 1, 1)
        # drivable raster of every agent's own frame: pixel (col, row) -> agent metres -> world -> road test
        col, row = torch.arange(224, device=d).float(), torch.arange(224, device=d).float()
        ax, ay = (col - 56.0) / 2.0, (row - 112.0) / 2.0
        loc = torch.stack(torch.broadcast_tensors(ax[None, :], ay[:, None]), -1)                      # [224(row), 224(col), 2]
        world = torch.einsum('bij,rcj->brci', wfa[:, :2, :2], loc) + wfa[:, None, None, :2, 2]
        dmap = self.drivable_world(world)
        # the other agents' planned futures (their last chosen plans) in the ego frame
        pw = self.plan_world.view(S, A, T, 2)
        So = max(A - 1, 1)
        others = torch.zeros(S, A, So, T, 2, device=d)
        for a in range(A):
            idx = [j for j in range(A) if j != a] or [a]
            others[:, a] = pw[:, idx]
        others = self._to_agent(others.view(B, So, T, 2), wfa)
        avail = torch.ones(B, So, T, dtype=torch.bool, device=d) if A > 1 else torch.zeros(B, So, T, dtype=torch.bool, device=d)
        return {
            'history_positions': self._to_agent(self.hist_pos, wfa),
            'history_yaws': (self.hist_yaw - self.yaw[:, None])[..., None],
            'history_availabilities': torch.ones(B, self.H, dtype=torch.bool, device=d),
            'curr_speed': self.speed.clone(),
            'extent': self.extent,
            'world_from_agent': wfa,
            'agent_from_world': torch.linalg.inv(wfa),
            'raster_from_agent': rfa,
            'scene_index': torch.arange(S, device=d).repeat_interleave(A),
            'drivable_map': dmap,
            'all_other_agents_future_positions': others.contiguous(),
            'all_other_agents_future_availability': avail,
            'target_pos': torch.stack([self.speed * T * self.dt, torch.zeros(B, device=d)], 1),
        }

    def curr_states(self):
        """What ContextEncoder.forward returns beside cond_feat: (x, y, v, yaw) of the current state in the agent frame."""
        B = self.S * self.A
        z = torch.zeros(B, device=self.device)
        return torch.stack([z, z, self.speed, z], 1)

    # ------------------------------------------------------------------ step
    def step(self, action, num_steps_to_take=None):
        """action: {'positions' [B,T,2], 'yaws' [B,T,1]} in the frame the observation was taken in; the agents follow it for
        `num_steps_to_take` steps (the simulator's n_step_action), then replan."""
        n = min(int(num_steps_to_take or self.n_step_action), self.T, self.num_steps - self.t)
        wfa = self._world_from_agent()
        R, p = wfa[:, :2, :2], wfa[:, :2, 2]
        pos_w = torch.einsum('bij,btj->bti', R, action['positions']) + p[:, None, :]                   # [B,T,2]
        yaw_w = action['yaws'][..., 0] + self.yaw[:, None]
        stationary = (action['positions'].abs().sum(dim=(1, 2)) == 0)                                  # disable_control_on_stationary
        pos_w = torch.where(stationary[:, None, None], self.pos[:, None, :].expand_as(pos_w), pos_w)
        yaw_w = torch.where(stationary[:, None], self.yaw[:, None].expand_as(yaw_w), yaw_w)
        self.plan_world = pos_w
        prev = torch.cat([self.pos[:, None, :], pos_w[:, :n - 1]], 1)
        # statistics of the executed steps (the rollout's failure indicators): off-road centre, centre distance < 0.8 m within a scene
        self.offroad_steps += (~self.drivable_world(pos_w[:, :n])).float().sum(1)
        pw = pos_w[:, :n].view(self.S, self.A, n, 2)
        dist = (pw[:, :, None] - pw[:, None, :]).norm(dim=-1)                                           # [S,A,A,n]
        eye = torch.eye(self.A, dtype=torch.bool, device=self.device)[None, :, :, None]
        self.collision_steps += ((dist < 0.8) & ~eye).any(dim=2).float().sum(-1).view(-1)
        self.hist_pos = torch.cat([self.hist_pos, pos_w[:, :n]], 1)[:, -self.H:]
        self.hist_yaw = torch.cat([self.hist_yaw, yaw_w[:, :n]], 1)[:, -self.H:]
        self.speed = torch.where(stationary, torch.zeros_like(self.speed), (pos_w[:, n - 1] - prev[:, n - 1]).norm(dim=-1) / self.dt)
        self.pos, self.yaw = pos_w[:, n - 1].clone(), yaw_w[:, n - 1].clone()
        self.t += n
        return n

    def metrics(self):
        steps = max(self.t, 1)
        return {'offroad_rate': float((self.offroad_steps / steps).mean()), 'collision_rate': float((self.collision_steps / steps).mean()),
                'steps': self.t}


def closed_loop_rollout(env, policy, context_fn, *, num_action_samples=1, sampler="ddpm", log=None, **policy_kw):
    """rollout.py:95-100.  `context_fn(obs) -> aux_info {'cond_feat', 'curr_states'}` is the context encoder in front of the policy
    (`vae.pre_vae` in the reference).  -> list of per-replan records {t, act_idx, guide_losses}."""
    records = []
    done = env.is_done()
    step_index = 0
    while not done:
        obs = env.get_observation()
        action, info = policy.get_action(obs, num_action_samples=num_action_samples, step_index=step_index, sampler=sampler,
                                         aux_info=context_fn(obs), **policy_kw)
        taken = env.step(action)
        records.append({'t': env.t, 'steps_taken': taken, 'act_idx': info['act_idx'],
                        'guide_losses': {k: v.mean().item() for k, v in info['guide_losses'].items()}})
        if log:
            log(records[-1])
        done = env.is_done()
        step_index += 1
    return records
