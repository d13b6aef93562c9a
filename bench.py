#!/usr/bin/env python
"""bench.py -- guided scenarios/s of the CLD sampling hot path on B200 (contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision bf16|fp32]

A "step" = ONE pass of the hot path over one batch of synthetic nuScenes-shaped scenes:
    S scenes x A agents x N samples  ->  K_d denoising steps (denoiser + posterior update [+ guidance])
    -> LSTM decode + unicycle rollout -> off-road / collision indicators [-> all-gather over ranks].
Workload = BASELINE.json configs[1]: 256 scenes x 16 agents, 50 denoising steps (n_timesteps=100, stride 2),
collision + off-road guidance on every intermediate step.  Weak scaling: every rank runs its own 256 scenes.

Printed JSON (rank 0, one line): metric/value/unit/... per the contract, plus
  roofline      dominant kernel (denoiser forward): algorithmic FLOPs / CUDA-event time vs measured bf16 peak
  cpu_baseline  the oracle (CPU port of the reference's PyTorch sampler) on a bounded sample of the workload
  e2e           same metric through the public API (cld_b200.DmModel.forward) from pinned HOST buffers
  gpu_launches  kernels launched by libcld_b200.so in the timed region
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_ROW_STEP = {52: 119.23e6, 104: 237.42e6}     # SURVEY.md App. B (hook-counted on the reference module)
WORKLOAD = dict(scenes=256, agents=16, samples=1, horizon=52, n_timesteps=100, stride=2)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE denoiser launch from the committed `ncu --set full` capture
# (profiles/r01_unet_tc_ncu_full_4096rows.csv: 274.22 MB + 36.65 MB); only known for the captured shape
NCU_DRAM_BYTES_PER_LAUNCH = {("bf16", 4096): 274220800 + 36645120}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CLD_BENCH_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--sampler", default="ddim", choices=["ddim", "ddpm"])
    ap.add_argument("--no-guidance", action="store_true")
    ap.add_argument("--scenes", type=int, default=WORKLOAD["scenes"])
    ap.add_argument("--cpu-scenes", type=int, default=1, help="scenes in the bounded CPU-baseline sample")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-lanes", action="store_true", help="do not time the 2-lane variant after the headline")
    ap.add_argument("--skip-context", action="store_true", help="do not time the context encoder (row a14) after the headline")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1393.0), d.get("hbm_gbs", 6538.3), "measured"
    return 1590.0, 6650.0, "fallback"


def cpu_oracle_rate(a, n_scenes, steps, warmup):
    """The oracle (CPU fp32 PyTorch restatement of the reference sampler, pinned against the real
    reference in oracle/make_golden.py) on `n_scenes` scenes of the same workload, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import cld_oracle as O
    from cld_b200 import default_algo_config, make_scenes
    from cld_b200.dm_model import DmModel
    from cld_b200.vae import VaeModel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    A, N, T = WORKLOAD["agents"], WORKLOAD["samples"], WORKLOAD["horizon"]
    algo = default_algo_config()
    torch.manual_seed(0)
    dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=WORKLOAD["n_timesteps"])
    vae = VaeModel(algo)
    unet_sd = {k: v.detach() for k, v in dm.model.state_dict().items()}
    dec_sd = {k: v.detach() for k, v in vae.lstmvae.lstm_dec.state_dict().items()}
    sched = O.make_schedule(WORKLOAD["n_timesteps"])
    aux, batch = make_scenes(n_scenes, A, horizon=T, seed=123, dense=True)
    R = n_scenes * A * N
    torch.manual_seed(7)
    x_init = torch.randn(R, T, 4)
    K = len(O.step_indices(WORKLOAD["n_timesteps"], WORKLOAD["stride"]))
    noises = torch.randn(K, R, T, 4) if a.sampler == "ddpm" else None
    gd = None if a.no_guidance else dict(dec_sd=dec_sd, cond=aux["cond_feat"], curr=aux["curr_states"], batch=batch,
                                         A=A, N=N, cfg=O.DEFAULT_GUIDANCE)

    def one():
        with torch.no_grad():
            out = O.sample(unet_sd, sched, aux["cond_feat"], x_init, noises, WORKLOAD["n_timesteps"],
                           WORKLOAD["stride"], a.sampler, guidance=gd)
            traj, _ = O.decode_rollout(dec_sd, out["pred_traj"], aux["cond_feat"], aux["curr_states"])
            O.indicators(traj[..., :2], batch)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    return n_scenes / dt, dt, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(a.steps, 3)), min(a.warmup, 1)
    rate, dt, cores = cpu_oracle_rate(a, a.cpu_scenes, steps, warm)
    sample = "%d of %d scenes x %d agents, %d denoising steps (%s)%s, oracle port of the reference's PyTorch sampler" % (
        a.cpu_scenes, a.scenes, WORKLOAD["agents"], 50, a.sampler, "" if a.no_guidance else " guided")
    print(json.dumps({
        "impl": "reference", "metric": "guided scenarios/sec (50-step DDIM)", "value": rate, "unit": "scenarios/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a),
        "cpu_baseline": {"value": rate, "unit": "scenarios/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "scenarios/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def context_encoder_rate(dev, agents=1024, iters=5, warmup=3, cpu_agents=256):
    """SURVEY.md sec. 8 row a14, measured beside the headline (NOT inside its timed region): ContextEncoder.forward
    (34 x 224 x 224 raster -> cond_feat) for `agents` agents through cld_b200.ContextEncoder / cld_context_forward, inputs resident
    in HBM, CUDA events.  Algorithmic FLOPs: 6.07 GFLOP per agent (hook-counted on the reference)."""
    import torch
    from cld_b200 import default_algo_config
    from cld_b200.context import ContextEncoder
    torch.manual_seed(0)
    ce = ContextEncoder(4, default_algo_config(), {"image": (34, 224, 224)}, max_agents=agents).to(dev)
    g = torch.Generator(device=dev).manual_seed(11)
    batch = {"image": (torch.rand(agents, 34, 224, 224, device=dev, generator=g) < 0.05).float(),
             "history_positions": torch.zeros(agents, 31, 2, device=dev), "history_yaws": torch.zeros(agents, 31, 1, device=dev),
             "curr_speed": torch.rand(agents, device=dev, generator=g) * 10}
    for _ in range(warmup):
        ce(batch)
    torch.cuda.synchronize()
    l0 = ce.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ce(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf_peak, _, peak_src = measured_peaks()
    tflops = 6.07e9 * agents / (ms * 1e-3) / 1e12
    out = {"value": agents / ms * 1e3, "unit": "agents/s", "agents": agents, "ms": ms, "iters": iters, "dtype": "bf16 (fp32 accumulate)",
           "algorithmic_gflop_per_agent": 6.07, "executed_gflop_per_agent": ce.conv_flops_per_agent() / 1e9,
           "achieved_tflops": tflops, "frac_of_bf16_peak": tflops / tf_peak, "peak_source": peak_src + " bf16_tflops_sustained",
           "gpu_launches": ce.launch_count() - l0, "bytes_in_per_agent": 34 * 224 * 224 * 4}
    # the same with the rasterisation fused in (cld_context_forward_history): map layers + history points of 16 agents per raster
    from cld_b200.synthetic import make_history_batch
    hb = make_history_batch(64, num_neighbors=15, seed=3)
    rep = (agents + 63) // 64
    maps = hb["maps"].to(dev).repeat(rep, 1, 1, 1)[:agents].contiguous()
    hpos = hb["agent_hist_pos"].to(dev).repeat(rep, 1, 1, 1)[:agents].contiguous()
    hmask = hb["agent_hist_mask"].to(dev).repeat(rep, 1, 1)[:agents].contiguous()
    b2 = dict(batch)
    b2["raster_from_agent"] = hb["raster_from_agent"].to(dev).repeat(rep, 1, 1)[:agents].contiguous()
    for _ in range(2):
        ce.forward_history(b2, maps, hpos, hmask)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ce.forward_history(b2, maps, hpos, hmask)
    e1.record()
    torch.cuda.synchronize()
    hms = e0.elapsed_time(e1) / iters
    out["from_history"] = {"value": agents / hms * 1e3, "unit": "agents/s", "ms": hms, "agents_per_raster": 16,
                           "bytes_in_per_agent": 3 * 224 * 224 * 4 + 16 * 31 * 9 + 36}
    # end to end through the public API from pinned HOST buffers: H2D of the un-rasterised inputs + forward_history + D2H of cond_feat
    host = {"maps": maps.cpu().pin_memory(), "hpos": hpos.cpu().pin_memory(), "hmask": hmask.cpu().pin_memory(),
            "rfa": b2["raster_from_agent"].cpu().pin_memory(), "history_positions": batch["history_positions"].cpu().pin_memory(),
            "history_yaws": batch["history_yaws"].cpu().pin_memory(), "curr_speed": batch["curr_speed"].cpu().pin_memory()}

    def e2e_step():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        o = ce.forward_history({"raster_from_agent": d["rfa"], "history_positions": d["history_positions"], "history_yaws": d["history_yaws"],
                                "curr_speed": d["curr_speed"]}, d["maps"], d["hpos"], d["hmask"])
        return o["cond_feat"].cpu()
    e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    ems = (time.perf_counter() - t0) / 3 * 1e3
    out["from_history"]["e2e"] = {"value": agents / ems * 1e3, "unit": "agents/s", "ms": ems,
                                  "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()), "d2h_bytes_per_step": agents * 256 * 4}
    ce.close()
    # the reference's CPU path for the same row, timed beside it: the oracle restatement of ContextEncoder.forward (bit-equal to the
    # real reference, oracle/make_golden.py) on a bounded sample with all host threads
    if cpu_agents:
        import cld_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        sd = {k: v.detach().cpu() for k, v in ce.state_dict().items()}
        sub = {k: v[:cpu_agents].cpu() for k, v in batch.items()}
        with torch.no_grad():
            O.context_encode(sd, {k: v[:2] for k, v in sub.items()})
            t0 = time.perf_counter()
            O.context_encode(sd, sub)
            dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cpu_agents / dt, "unit": "agents/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "%d of %d agents, 1 timed pass (%.1f s)" % (cpu_agents, agents, dt)}
    return out


def raster_chain_rate(dev, S, A, batch_d, hot_path, algo, iters=2):
    """The reference's inference chain in one timed region (rows a14 + a1..a13): rasters resident in HBM ->
    ContextEncoder.forward -> DmModel.forward (guided sampler + decode + rollout + indicators), same workload as the headline."""
    import torch
    from cld_b200.context import ContextEncoder
    B = S * A
    torch.manual_seed(0)
    ce = ContextEncoder(4, algo, {"image": (34, 224, 224)}, max_agents=B).to(dev)
    g = torch.Generator(device=dev).manual_seed(13)
    b2 = dict(batch_d)
    b2["image"] = (torch.rand(B, 34, 224, 224, device=dev, generator=g) < 0.05).float()
    b2["history_yaws"] = torch.zeros(B, 31, 1, device=dev)

    def step():
        aux2 = ce(b2)
        return hot_path(b2, {"cond_feat": aux2["cond_feat"], "curr_states": aux2["curr_states"]})
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ce.close()
    del b2
    return {"value": S / ms * 1e3, "unit": "scenarios/s", "ms_per_step": ms, "iters": iters, "raster_bytes_per_scene": A * 34 * 224 * 224 * 4}


def workload_config(a):
    return {"workload": "cfg1: %d scenes x %d agents x %d sample, T=%d, n_timesteps=%d stride %d (50 denoising steps, %s)%s, "
                        "decode+rollout+indicators; random-init weights" % (
                            a.scenes, WORKLOAD["agents"], WORKLOAD["samples"], WORKLOAD["horizon"], WORKLOAD["n_timesteps"],
                            WORKLOAD["stride"], a.sampler, "" if a.no_guidance else ", agent_collision+map_collision guidance"),
            "rows_per_gpu": a.scenes * WORKLOAD["agents"] * WORKLOAD["samples"], "precision": a.precision,
            "l2": "working set per denoising step exceeds L2 only for the guidance stash (545 MB); denoiser "
                  "activations are on-chip; no L2 flush needed between steps (inputs change every denoising step)",
            "parallelism": "scene-sharded, %d rank(s)" % a.gpus}


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cld_b200 import default_algo_config, make_scenes
    from cld_b200.dm_model import DmModel
    from cld_b200.engine import default_guidance
    from cld_b200.vae import VaeModel
    S, A, N, T = a.scenes, WORKLOAD["agents"], WORKLOAD["samples"], WORKLOAD["horizon"]
    R = S * A * N
    algo = default_algo_config(num_samp=N)
    torch.manual_seed(0)
    dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=WORKLOAD["n_timesteps"], precision=a.precision,
                 max_rows=R).to(dev)
    dm.stride = WORKLOAD["stride"]
    vae = VaeModel(algo).bind(dm)
    aux, batch = make_scenes(S, A, horizon=T, seed=123 + rank, dense=True)
    guidance = None if a.no_guidance else default_guidance()
    eng = dm.engine(R)
    K_d = len(range(0, WORKLOAD["n_timesteps"], WORKLOAD["stride"]))

    # ---- device-resident inputs (for `value`) and pinned host copies (for `e2e`)
    aux_d = {k: v.to(dev) for k, v in aux.items()}
    batch_d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    host_keys = ["extent", "world_from_agent", "raster_from_agent", "curr_speed", "drivable_map", "scene_index",
                 "all_other_agents_future_positions", "all_other_agents_future_availability", "history_positions"]
    aux_h = {k: v.pin_memory() for k, v in aux.items()}
    batch_h = {k: batch[k].pin_memory() for k in host_keys}
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    x_init = torch.randn(R, T, 4, device=dev, generator=gen)
    noise = torch.randn(K_d, R, T, 4, device=dev, generator=gen) if a.sampler == "ddpm" else None
    from cld_b200.distributed import gather_results
    gathered = None
    if world > 1:
        gathered = torch.empty(world * R, T * 6 + T + 1, device=dev)

    def hot_path(b, ax):
        out = dm(b, ax, algo, x_init=x_init, noise=noise, sampler=a.sampler, guidance=guidance, want_indicators=True,
                 agents_per_scene=A)
        if world > 1:
            # the path's one exchange: trajectories + indicator flags + collision counts of every rank
            out["all_traj"], out["all_offroad"], out["all_coll"] = gather_results(
                out["traj"], out["offroad"], out["coll"], out=gathered)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        hot_path(batch_d, aux_d)
    barrier()
    sampler_thread = ClockSampler(local) if rank == 0 else None
    if sampler_thread:
        sampler_thread.start()
    l0 = eng.launch_count()
    eng.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = hot_path(batch_d, aux_d)
    e1.record()
    barrier()
    prof = eng.profile_end()
    launches = eng.launch_count() - l0
    clocks = sampler_thread.stop() if sampler_thread else None
    ms = e0.elapsed_time(e1) / a.steps
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = tms.item()
    value = world * S / (ms / 1e3)

    # ---- e2e: public API from pinned host buffers, H2D inputs + D2H results inside the timed region
    def e2e_step():
        b = {k: v.to(dev, non_blocking=True) for k, v in batch_h.items()}
        ax = {k: v.to(dev, non_blocking=True) for k, v in aux_h.items()}
        o = hot_path(b, ax)
        res = (o["traj"].cpu(), o["offroad"].cpu(), o["coll"].cpu())
        return res
    h2d = sum(v.numel() * v.element_size() for v in batch_h.values()) + sum(v.numel() * v.element_size() for v in aux_h.values())
    d2h = R * T * 6 * 4 + R * T + R * 4
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(a.steps, 3))
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) / n_e2e * 1e3
    te = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * S / (te.item() / 1e3)

    if rank == 0:
        tf_peak, hbm_peak, peak_src = measured_peaks()
        den_ms, den_n = prof["denoiser"]
        per_launch_ms = den_ms / max(den_n, 1)
        flop = FLOP_PER_ROW_STEP[T] * R
        achieved = flop / (per_launch_ms * 1e-3) / 1e12 if den_n else 0.0
        roof = {"bound": "tensor", "kernel": "denoiser forward (%s)" % ("unet_tc megakernel" if a.precision == "bf16" else "fp32 SIMT layer kernels"),
                "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((a.precision, R)), "traffic_unit": "bytes (dram read + write per launch)",
                "traffic_source": "profiles/r01_unet_tc_ncu_full_4096rows.csv (ncu --set full, one launch)",
                "peak_source": peak_src + " bf16_tflops_sustained", "per_launch_ms": per_launch_ms, "launches_timed": den_n,
                "algorithmic_flop_per_launch": flop,
                "share_of_step": {k: v[0] / a.steps / ms for k, v in prof.items()}}
        cpu = None
        if not a.skip_cpu and world == 1:
            rate, dt, cores = cpu_oracle_rate(a, a.cpu_scenes, 1, 0)
            cpu = {"value": rate, "unit": "scenarios/s", "cores": cores, "kind": "port",
                   "sample": "%d of %d scenes x %d agents, same 50-step %s%s sampler + decode + indicators, 1 timed pass (%.1f s)" % (
                       a.cpu_scenes, S, A, a.sampler, "" if a.no_guidance else " guided", dt)}
        lanes = None
        if world == 1 and not a.skip_lanes:
            # informational: the same workload with DmModel(lanes=2) -- whole-scene half batches on two engines / CUDA streams fill the
            # SMs that the denoiser's last wave and the 128-CTA decoder kernels leave idle.  Not the headline: concurrent lanes blur the
            # per-launch timing the roofline is quoted on.
            torch.manual_seed(0)
            dm2 = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=WORKLOAD["n_timesteps"], precision=a.precision, max_rows=R,
                          lanes=2).to(dev)
            dm2.stride = WORKLOAD["stride"]
            VaeModel(algo).bind(dm2)

            def lane_step():
                return dm2(batch_d, aux_d, algo, x_init=x_init, noise=noise, sampler=a.sampler, guidance=guidance, want_indicators=True,
                           agents_per_scene=A)
            for _ in range(2):
                lane_step()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(a.steps):
                lane_step()
            f1.record()
            torch.cuda.synchronize()
            lms = f0.elapsed_time(f1) / a.steps
            lanes = {"lanes": 2, "value": S / lms * 1e3, "unit": "scenarios/s", "ms_per_step": lms}
            del dm2
        ctx = None
        if world == 1 and not a.skip_context:
            chain = raster_chain_rate(dev, S, A, batch_d, hot_path, algo)
            del out, x_init, noise
            torch.cuda.empty_cache()
            ctx = context_encoder_rate(dev, cpu_agents=0 if a.skip_cpu else 256)
            ctx["raster_to_trajectories"] = chain
        print(json.dumps({
            "metric": "guided scenarios/sec (50-step DDIM)", "value": value, "unit": "scenarios/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.precision, "data": "synthetic", "config": workload_config(a),
            "row_steps_per_s": value * A * N * K_d, "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "scenarios/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": te.item()},
            "gpu_launches": launches, "clocks": clocks, "context_encoder": ctx, "lanes": lanes,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
