"""Multi-GPU (torchrun): the sharded InfoNCE op (queue K/G per rank) with the exchange done by this repo's kernels over
NVLink peer memory ("peer": prep stores the packed queries into every rank's table, reduce_scatter_kernel stores the
partial results into the row owners' accumulators, two device barriers) against NCCL all_gather + reduce_scatter
("nccl").  CUDA events on each rank, max over ranks.

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/exchange_bench.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.nn.functional as F
from mscl_b200 import functional as fx

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
rows = []
for K in (65536, 1048576):
    nq = fx.NegativeQueue(K, 128, dev, rank, world, shard=True)
    g = torch.Generator().manual_seed(0)
    nq.load(F.normalize(torch.randn(128, K, generator=g), dim=0), torch.randint(0, 100, (K,), generator=g), 0)
    for M in (32, 96):
        g = torch.Generator().manual_seed(rank)
        q = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev).requires_grad_(True)
        k = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev)
        res = dict(K=K, M_local=M, world=world)
        for mode in ("peer", "nccl"):
            fx.EXCHANGE = mode
            for _ in range(5):
                out, _ = fx.infonce(q, k, nq, M, 0.07, group=dist.group.WORLD)
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 50
            e0.record()
            for _ in range(n):
                out, _ = fx.infonce(q, k, nq, M, 0.07, group=dist.group.WORLD)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) * 1e3 / n], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[mode + "_us"] = float(t)
            res[mode + "_loss"] = float(out[0, 0])
        rows.append(res)
        if rank == 0:
            print(f"G={world} K={K} (K/G={K // world}) M/rank={M}: peer-memory exchange {res['peer_us']:.1f} us/op, "
                  f"NCCL exchange {res['nccl_us']:.1f} us/op  (loss {res['peer_loss']:.5f} / {res['nccl_loss']:.5f})", flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(f"gpurun_out/exchange_bench_g{world}.json", "w"), indent=1)
dist.destroy_process_group()
