"""Timing of the context encoder (cld_context_forward) on one B200, next to PyTorch eager (cuDNN, bf16 channels-last and
fp32) running the oracle's ops on the same GPU.  Usage: python tools/time_context.py [agents] [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import cld_oracle as O
from cld_b200 import default_algo_config
from cld_b200.context import ContextEncoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
eager = "--no-eager" not in sys.argv
torch.manual_seed(0)
ce = ContextEncoder(4, default_algo_config(), {"image": (34, 224, 224)}, max_agents=B)
sd = O.synth_context_state({k: tuple(v.shape) for k, v in ce.state_dict().items()})
ce.load_state_dict(sd)
ce = ce.cuda()
g = torch.Generator(device="cuda").manual_seed(1)
img = (torch.rand(B, 34, 224, 224, device="cuda", generator=g) < 0.05).float()
batch = {"image": img, "history_positions": torch.zeros(B, 31, 2, device="cuda"), "history_yaws": torch.zeros(B, 31, 1, device="cuda"),
         "curr_speed": torch.rand(B, device="cuda") * 10}


def timeit(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timeit(lambda: ce(batch), iters)
fl = ce.conv_flops_per_agent()
print("cld_context_forward: B=%d  %.2f ms  %.0f agents/s  executed conv %.2f GFLOP/agent -> %.1f TFLOP/s (algorithmic 6.07 GFLOP/agent -> %.1f TFLOP/s)"
      % (B, ms, B / ms * 1e3, fl / 1e9, fl * B / ms / 1e9, 6.07e9 * B / ms / 1e9))
# fused history rasteriser: map layers + history points instead of the fp32 image
from cld_b200.synthetic import make_history_batch
hb = make_history_batch(64, num_neighbors=15, seed=3)
rep = (B + 63) // 64
maps = hb["maps"].cuda().repeat(rep, 1, 1, 1)[:B].contiguous()
hpos = hb["agent_hist_pos"].cuda().repeat(rep, 1, 1, 1)[:B].contiguous()
hmask = hb["agent_hist_mask"].cuda().repeat(rep, 1, 1)[:B].contiguous()
b2 = dict(batch)
b2["raster_from_agent"] = hb["raster_from_agent"].cuda().repeat(rep, 1, 1)[:B].contiguous()
ms2 = timeit(lambda: ce.forward_history(b2, maps, hpos, hmask), iters)
print("cld_context_forward_history (16 agents per raster): B=%d  %.2f ms  %.0f agents/s" % (B, ms2, B / ms2 * 1e3))
if eager:
    sdc = {k: v.cuda() for k, v in sd.items()}
    Bs = min(B, 512)
    sub = {k: v[:Bs] for k, v in batch.items()}
    with torch.no_grad():
        ms32 = timeit(lambda: O.context_encode(sdc, sub), 3)
        sd16 = {k: (v.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if v.dim() == 4 else (v.to(torch.bfloat16) if v.is_floating_point() else v)) for k, v in sdc.items()}
        sub16 = dict(sub)

        def run16():
            b = dict(sub16)
            b["image"] = sub["image"].to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            b["history_positions"] = sub["history_positions"].to(torch.bfloat16)
            b["history_yaws"] = sub["history_yaws"].to(torch.bfloat16)
            b["curr_speed"] = sub["curr_speed"].to(torch.bfloat16)
            return O.context_encode(sd16, b)
        ms16 = timeit(run16, 3)
    print("PyTorch eager on the same GPU (B=%d): fp32 %.2f ms (%.0f agents/s) | bf16 channels-last %.2f ms (%.0f agents/s)"
          % (Bs, ms32, Bs / ms32 * 1e3, ms16, Bs / ms16 * 1e3))
