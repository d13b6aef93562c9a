"""Drop-in for the reference's ContextEncoder (models/context_utils.py:8-61): the provider of `cond_feat` /
`curr_states` for the sampler.

Same constructor, same `forward(data_batch) -> {'cond_feat', 'curr_states', 'image'}`, same state-dict keys
(`agent_state_encoder._model.*`, `map_encoder.encoder_heads.map_model.*`, `process_cond_mlp._model.*`; a reference
`vae.context_encoder.*` checkpoint loads unchanged).  The modules below only HOLD parameters; all arithmetic runs in
libcld_b200.so (`cld_context_forward`: tcgen05 implicit-GEMM ResNet-18 + fused MLP head).  No CPU / PyTorch fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from ._lib import lib


class _MLP(nn.Module):
    """Parameter container of base_models.MLP(normalization=True) (src/tbsim/models/base_models.py:58-66):
    `_model` = Sequential(Linear, LayerNorm, ReLU, ..., Linear)."""

    def __init__(self, input_dim, output_dim, layer_dims):
        super().__init__()
        layers, dim = [], input_dim
        for l in layer_dims:
            layers += [nn.Linear(dim, l), nn.LayerNorm(l), nn.ReLU()]
            dim = l
        layers.append(nn.Linear(dim, output_dim))
        self._model = nn.Sequential(*layers)


class _BasicBlock(nn.Module):
    """torchvision.models.resnet.BasicBlock parameter layout (conv1, bn1, conv2, bn2[, downsample.{0,1}])."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class _ResNet18(nn.Module):
    """RasterizedMapEncoder.map_model (base_models.py:573-607): resnet18 with a `num_input_channels` 7x7 stem and
    fc 512 -> feature_dim."""

    def __init__(self, in_channels, feature_dim):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, cout in enumerate((64, 128, 256, 512), start=1):
            stride = 1 if li == 1 else 2
            setattr(self, "layer%d" % li, nn.Sequential(_BasicBlock(cin, cout, stride), _BasicBlock(cout, cout, 1)))
            cin = cout
        self.fc = nn.Linear(512, feature_dim)


class _Holder(nn.Module):
    pass


class ContextEncoder(nn.Module):
    def __init__(self, state_in_dim, algo_config, modality_shapes, dyn=None, *, max_agents=None):
        super().__init__()
        self.dyn = dyn
        d = algo_config.curr_state_feat_dim
        if (d, algo_config.map_feature_dim, algo_config.cond_feat_dim) != (64, 256, 256) or state_in_dim != 4:
            raise RuntimeError("the B200 context encoder is built for curr_state_feat_dim 64, map_feature_dim 256, "
                               "cond_feat_dim 256 and the 4-d unicycle state (the reference's config.yaml)")
        if algo_config.map_encoder_model_arch != "resnet18" or tuple(modality_shapes["image"]) != (34, 224, 224):
            raise RuntimeError("the B200 context encoder implements resnet18 on a 34 x 224 x 224 raster")
        self.agent_state_encoder = _MLP(state_in_dim, d, (d, d))
        self.map_encoder = _Holder()
        self.map_encoder.encoder_heads = _Holder()
        self.map_encoder.encoder_heads.map_model = _ResNet18(modality_shapes["image"][0], algo_config.map_feature_dim)
        cin = d + algo_config.map_feature_dim
        cout = algo_config.cond_feat_dim
        self.process_cond_mlp = _MLP(cin, cout, (cin, cin, cout, cout))
        # workspace capacity in agents (5.9 MB each, processed in chunks of <= 2048); None: sized by the largest batch seen
        self._max_agents = None if max_agents is None else int(max_agents)
        self._capacity = 0
        self._handle = None
        self._handle_dev = None
        self._dirty = True
        self.eval()

    # ------------------------------------------------------------------ engine plumbing
    def invalidate(self):
        """Force a re-pack of the parameters at the next call (normally not needed, see `_weights_signature`)."""
        self._dirty = True

    def _weights_signature(self):
        """(identity, in-place version) of every parameter / buffer: changes on any load_state_dict (also a parent module's,
        which never calls this module's own load_state_dict), optimizer step or manual copy_."""
        return hash(tuple((id(v), v._version) for v in list(self.parameters()) + list(self.buffers())))

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._dirty = True
        return out

    def close(self):
        if self._handle is not None:
            lib.cld_context_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, what):
        msg = lib.cld_context_last_error(self._handle if self._handle is not None else C.c_void_p(0))
        raise RuntimeError("%s: %s" % (what, (msg or b"").decode()))

    def weight_list(self):
        """The 130 fp32 tensors `cld_context_load` takes (state-dict order without num_batches_tracked)."""
        return [v for k, v in self.state_dict().items() if not k.endswith("num_batches_tracked")]

    def _engine(self, B):
        p = next(self.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("cld_b200.ContextEncoder runs on a B200 only (module is on %s); there is no CPU fallback" % p.device)
        need = max(self._max_agents, 1) if self._max_agents is not None else min(max(int(B), 1), 2048)
        if self._handle is None or self._handle_dev != p.device or (self._max_agents is None and need > self._capacity):
            self.close()
            with torch.cuda.device(p.device):
                h = C.c_void_p()
                if lib.cld_context_create(int(need), C.byref(h)) != 0:
                    self._handle = None
                    self._err("cld_context_create")
            self._handle, self._handle_dev, self._dirty, self._capacity = h, p.device, True, need
        sig = self._weights_signature()
        if self._dirty or sig != getattr(self, "_loaded_sig", None):
            self._loaded_sig = sig
            ws = [w.detach().to(torch.float32).contiguous() for w in self.weight_list()]
            ptrs = (C.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
            numels = (C.c_int64 * len(ws))(*[w.numel() for w in ws])
            stream = torch.cuda.current_stream(p.device).cuda_stream
            if lib.cld_context_load(self._handle, ptrs, numels, len(ws), C.c_void_p(stream)) != 0:
                self._err("cld_context_load")
            torch.cuda.current_stream(p.device).synchronize()     # `ws` may hold temporaries
            self._dirty = False
        return self._handle

    @staticmethod
    def current_states(data_batch):
        """batch_utils.get_current_states, unicycle branch (src/tbsim/utils/batch_utils.py:61-65): [x, y, vel, yaw]."""
        return torch.cat([data_batch["history_positions"][..., -1, :], data_batch["curr_speed"][..., None],
                          data_batch["history_yaws"][..., -1, :1]], dim=-1)

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def forward(self, data_batch, *, tap_stage=None, want_map_feat=False):
        image = data_batch["image"]
        dev = image.device
        B = image.shape[0]
        h = self._engine(B)
        if dev != self._handle_dev:
            raise RuntimeError("ContextEncoder: data_batch['image'] is on %s but the module is on %s" % (dev, self._handle_dev))
        if tuple(image.shape[1:]) != (34, 224, 224):
            raise RuntimeError("ContextEncoder: expected a [B,34,224,224] raster, got %s" % (tuple(image.shape),))
        image = image.to(torch.float32).contiguous()
        curr = self.current_states(data_batch).to(device=dev, dtype=torch.float32).contiguous()
        cond = torch.empty(B, 256, device=dev, dtype=torch.float32)
        map_feat = torch.empty(B, 256, device=dev, dtype=torch.float32) if want_map_feat else None
        tap = None
        if tap_stage is not None:
            hw, ch = {0: (56, 64), 1: (56, 64), 2: (28, 128), 3: (14, 256), 4: (7, 512)}[int(tap_stage)]
            tap = torch.empty(B, ch, hw, hw, device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.cld_context_forward(h, C.c_void_p(image.data_ptr()), C.c_void_p(curr.data_ptr()), B, C.c_void_p(cond.data_ptr()),
                                     C.c_void_p(map_feat.data_ptr() if map_feat is not None else 0),
                                     -1 if tap_stage is None else int(tap_stage), C.c_void_p(tap.data_ptr() if tap is not None else 0),
                                     C.c_void_p(stream))
        if rc != 0:
            self._err("cld_context_forward")
        out = {"cond_feat": cond, "curr_states": curr, "image": data_batch["image"]}
        if map_feat is not None:
            out["map_feat"] = map_feat
        if tap is not None:
            out["tap"] = tap
        return out

    def _history_args(self, maps, agent_hist_pos, agent_hist_mask, raster_from_agent):
        dev = maps.device
        B, A, T = agent_hist_pos.shape[:3]
        if tuple(maps.shape) != (B, 3, 224, 224) or T != 31 or tuple(agent_hist_mask.shape) != (B, A, T):
            raise RuntimeError("history rasteriser: expected maps [B,3,224,224], positions [B,A,31,2], mask [B,A,31]; got %s %s %s"
                               % (tuple(maps.shape), tuple(agent_hist_pos.shape), tuple(agent_hist_mask.shape)))
        return (maps.to(torch.float32).contiguous(), agent_hist_pos.to(device=dev, dtype=torch.float32).contiguous(),
                agent_hist_mask.to(device=dev).to(torch.uint8).contiguous(),
                raster_from_agent.to(device=dev, dtype=torch.float32).contiguous(), B, A)

    @torch.no_grad()
    def rasterize_agents(self, maps, agent_hist_pos, agent_hist_yaw, agent_mask, raster_from_agent, map_res=None):
        """Drop-in for tbsim.utils.trajdata_utils.rasterize_agents (src/tbsim/utils/trajdata_utils.py:123-156): returns the
        [B,34,224,224] fp32 image (`agent_hist_yaw` / `map_res` are unused there as well)."""
        self._engine(maps.shape[0])
        m, p, k, r, B, A = self._history_args(maps, agent_hist_pos, agent_mask, raster_from_agent)
        img = torch.empty(B, 34, 224, 224, device=m.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(m.device).cuda_stream
        rc = lib.cld_context_forward_history(self._handle, C.c_void_p(m.data_ptr()), C.c_void_p(p.data_ptr()), C.c_void_p(k.data_ptr()),
                                             C.c_void_p(r.data_ptr()), A, None, B, None, None, C.c_void_p(img.data_ptr()), C.c_void_p(stream))
        if rc != 0:
            self._err("cld_context_forward_history")
        return img

    @torch.no_grad()
    def forward_history(self, data_batch, maps, agent_hist_pos, agent_hist_mask, *, want_image=False):
        """ContextEncoder.forward with the rasterisation fused in: `maps` [B,3,224,224] map layers, `agent_hist_pos` [B,A,31,2]
        (ego first) and `agent_hist_mask` [B,A,31] are what parse_node_centric hands to rasterize_agents
        (trajdata_utils.py:395-420); `data_batch` supplies raster_from_agent, history_positions / history_yaws / curr_speed."""
        self._engine(maps.shape[0])
        m, p, k, r, B, A = self._history_args(maps, agent_hist_pos, agent_hist_mask, data_batch["raster_from_agent"])
        dev = m.device
        curr = self.current_states(data_batch).to(device=dev, dtype=torch.float32).contiguous()
        cond = torch.empty(B, 256, device=dev, dtype=torch.float32)
        img = torch.empty(B, 34, 224, 224, device=dev, dtype=torch.float32) if want_image else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.cld_context_forward_history(self._handle, C.c_void_p(m.data_ptr()), C.c_void_p(p.data_ptr()), C.c_void_p(k.data_ptr()),
                                             C.c_void_p(r.data_ptr()), A, C.c_void_p(curr.data_ptr()), B, C.c_void_p(cond.data_ptr()), None,
                                             C.c_void_p(img.data_ptr()) if img is not None else None, C.c_void_p(stream))
        if rc != 0:
            self._err("cld_context_forward_history")
        out = {"cond_feat": cond, "curr_states": curr}
        if img is not None:
            out["image"] = img
        return out

    def launch_count(self):
        return int(lib.cld_context_launch_count(self._handle)) if self._handle is not None else 0

    def conv_flops_per_agent(self):
        return float(lib.cld_context_conv_flops(self._handle)) if self._handle is not None else 0.0
