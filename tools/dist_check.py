"""Run under torchrun on N GPUs: frame-range sharded synthesis (NCCL all-gather of OLA halos) must equal
the unsharded single-GPU result; prints one line per rank."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import libllsm2_b200 as L
from libllsm2_b200 import parallel
from libllsm2_b200.synthetic import synth_frames

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, F = 8, 400 * world                                     # long utterances: F frames split over the ranks
fr, conf = synth_frames(B, F, seed=7, nhar=64, maxnhar=64)
dev = torch.device("cuda", local)
d = {k: (torch.from_numpy(v).to(dev) if v is not None else None) for k, v in fr.items()}
ctx = L.Context(local)
lo, hi = parallel.frame_shards(F, world)[rank]
part = L.synthesize_l0_shard(ctx, conf, d, lo, hi, white=None, seed=99)
ex = parallel.HaloExchange(ctx, conf, world, rank, part["y"].shape[1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sa, sb = ex.exchange(part)                  # llsm_b200_halo_exchange: pack kernel -> ncclAllGather -> edge-add kernel
e1.record(); torch.cuda.synchronize()
out = {k: v[:, sa:sb] for k, v in part.items()}
full = L.synthesize_l0(ctx, conf, d, white=None, seed=99)      # every rank also computes the whole thing
torch.cuda.synchronize()
err = {k: float((out[k] - full[k][:, sa:sb]).abs().max()) for k in ("y", "y_sin", "y_noise")}
print("rank %d/%d frames [%d,%d) samples [%d,%d) halo-exchange %.3f ms  max|sharded - unsharded| %s  rms(y)=%.4f"
      % (rank, world, lo, hi, sa, sb, e0.elapsed_time(e1), err, float(full["y"].pow(2).mean().sqrt())), flush=True)
assert all(v < 1e-5 for v in err.values()), err
dist.barrier()
dist.destroy_process_group()
