"""2-rank NCCL check that a sharded run equals the unsharded one (launched by tests/test_gpu_parity.py under torchrun):
rank r samples the scenes `shard_scenes` gives it with its global row offset, one all-gather, rank 0 compares with the whole
batch sampled on its own GPU.  Prints SHARD_CHECK_OK."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cld_b200 import default_algo_config, make_scenes                      # noqa: E402
from cld_b200.distributed import gather_results, shard_batch, shard_scenes   # noqa: E402
from cld_b200.dm_model import DmModel                                       # noqa: E402
from cld_b200.engine import default_guidance                                # noqa: E402
from cld_b200.vae import VaeModel                                           # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    S, A, N = 8, 16, 1
    algo = default_algo_config(num_samp=N)
    torch.manual_seed(0)
    dm = DmModel(algo, {"image": (34, 224, 224)}, n_timesteps=10, precision="bf16", max_rows=S * A * N).to(dev)
    VaeModel(algo).bind(dm)
    aux, batch = make_scenes(S, A, seed=77, dense=True)
    torch.manual_seed(5)
    x_init = torch.randn(S * A * N, 52, 4)
    kw = dict(sampler="ddpm", guidance=default_guidance(), use_device_rng=True, seed=31337, want_indicators=True, agents_per_scene=A)
    cu = lambda d: {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}   # noqa: E731
    s0, s1 = shard_scenes(S, world, rank)
    r0, r1 = s0 * A * N, s1 * A * N
    mine = dm(cu(shard_batch(batch, A, world, rank)), cu(shard_batch(aux, A, world, rank)), algo, x_init=x_init[r0:r1].to(dev),
              row_offset=r0, **kw)
    traj, off, coll = gather_results(mine["traj"], mine["offroad"], mine["coll"])
    ok = True
    if rank == 0:
        full = dm(cu(batch), cu(aux), algo, x_init=x_init.to(dev), **kw)
        ok = torch.equal(traj, full["traj"]) and torch.equal(off, full["offroad"]) and torch.equal(coll, full["coll"])
        print("SHARD_CHECK_OK" if ok else "SHARD_CHECK_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
