"""torchrun --nproc-per-node N tools/p2p_check.py: PeerGather (the spectrum kernel stores straight into the
destination rank's window over NVLink) against the NCCL gather of evaluate_sharded on the same shards -- same bits on
the destination rank -- and the step time of both."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from relagn_b200.batch import BatchEvaluator, evaluate_sharded, PeerGather

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
ev = BatchEvaluator(bench.ebins(), model="relagn", device=local)
ev.ensure_tables()
params = torch.from_numpy(bench.c4_params(P, bench.SEED + rank)).to(dev)
out = torch.empty((P, ev.nE), dtype=torch.float64, device=dev)
_, g_nccl = evaluate_sharded(ev, params, rel=True, gather_dst=0, out=out)
g_nccl_full = None
if rank == 0:
    g_nccl_full = [torch.empty((P, ev.nE), dtype=torch.float64, device=dev) for _ in range(world)]
    for a, b in zip(g_nccl_full, g_nccl):
        a.copy_(b)
tr = PeerGather(P, ev.nE, dev)
for rep in range(3):  # three steps: both halves of the window and the reuse of the first
    _, g_p2p = evaluate_sharded(ev, params, rel=True, gather_dst=0, transport=tr)
    torch.cuda.synchronize()
    if rank == 0:
        same = all(torch.equal(a, b) for a, b in zip(g_nccl_full, g_p2p))
        print("step %d: window rows equal the nccl gather: %s" % (rep, same), flush=True)
# `sampler` / `attrs` as further arguments: the two things bench.py's step has that this check does not
sampler = bench.ClockSampler(local) if "sampler" in sys.argv[2:] and rank == 0 else None
if sampler is not None:
    sampler.start()
extra = {"attrs": torch.empty((P, 24), dtype=torch.float64, device=dev)} if "attrs" in sys.argv[2:] else {}
for name, kw in (("nccl pieces", {}), ("direct stores into the window", {"transport": tr})):
    for it in range(4):
        if it == 1:
            dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        evaluate_sharded(ev, params, rel=True, gather_dst=0, out=out if not kw else None, gathered=g_nccl_full if not kw else None, **extra, **kw)
    dist.barrier(); torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / 3], device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%s: %.1f ms per step, %.0f spectra/s" % (name, dt.item() * 1e3, P * world / dt.item()), flush=True)
if sampler is not None:
    sampler.stop()
dist.destroy_process_group()
