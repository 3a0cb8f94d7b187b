"""BASELINE.json configs[2] on G GPUs (torchrun): the clip-level InfoNCE op over K = 16 Ki ... 1 Mi negatives with the
queue SHARDED K/G per GPU (peer-memory exchange of the packed queries and of the partial results, csrc/infonce.cu +
csrc/infonce_fused.cu) against the same op with the queue REPLICATED on every GPU (one launch, csrc/infonce_fused.cu).

Per (K, M per GPU): us per op (CUDA events on every rank, max over ranks, launch trains), the algorithmic HBM rate per GPU
(functional.infonce_algo_bytes of what ONE GPU streams: its K/G shard for all G*M gathered rows, or the whole queue for
its own M rows), the fraction of the measured HBM peak, and the tensor-core rate of the all-gathered regime
(4 * M_all * K_local * 128 FLOP per GPU; tf32 dense peak 1.1 PFLOP/s nominal).

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/k1_sweep_multi.py [--out FILE]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.nn.functional as F
from mscl_b200 import functional as fx

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--Ks", default="16384,65536,262144,1048576")
ap.add_argument("--Ms", default="64,192")
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        peak = float(json.load(f)["hbm_gbs"])
except Exception:
    peak = 6650.0


def timed(fn, n, warm=4):
    for _ in range(warm):       # every queue of the ring once: the first sharded op on a queue builds its peer workspace
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


rows = []
for K in [int(x) for x in args.Ks.split(",")]:
    g = torch.Generator().manual_seed(K)
    qk = F.normalize(torch.randn(128, K, generator=g), dim=0)
    cnt = torch.randint(0, 2000, (K,), generator=g)
    # rings of queues so that consecutive launches do not find their tiles in L2 (2 x 126 MB)
    n_sh = max(2, min(24, -(-256_000_000 // (K // world * 512))))
    n_rep = max(2, min(24, -(-256_000_000 // (K * 512))))
    sharded = []
    for _ in range(n_sh):
        nq = fx.NegativeQueue(K, 128, dev, rank, world, shard=True)
        nq.load(qk, cnt, 0)
        sharded.append(nq)
    replicated = []
    for _ in range(n_rep):
        nq = fx.NegativeQueue(K, 128, dev)
        nq.load(qk, cnt, 0)
        replicated.append(nq)
    for M in [int(x) for x in args.Ms.split(",")]:
        g = torch.Generator().manual_seed(1000 + rank)
        q = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev).requires_grad_(True)
        k = F.normalize(torch.randn(M, 128, generator=g), dim=1).to(dev)
        st = dict(i=0)

        def op(queues):
            nq = queues[st["i"] % len(queues)]
            st["i"] += 1
            nq._fresh = False
            out, _ = fx.infonce(q, k, nq, M, 0.07, group=dist.group.WORLD)
            return out

        us_sh = timed(lambda: op(sharded), args.iters, warm=max(4, len(sharded) + 1))
        loss_sh = float(op(sharded)[0, 0])
        us_rep = timed(lambda: op(replicated), args.iters, warm=max(4, len(replicated) + 1))
        loss_rep = float(op(replicated)[0, 0])
        M_all, K_loc = M * world, K // world
        b_sh, b_rep = fx.infonce_algo_bytes(M_all, K_loc), fx.infonce_algo_bytes(M, K)
        row = dict(world=world, K=K, K_local=K_loc, M_local=M, M_all=M_all,
                   sharded_us=us_sh, sharded_gbs_per_gpu=b_sh / us_sh / 1e3, sharded_frac_hbm=b_sh / us_sh / 1e3 / peak,
                   sharded_tflops_per_gpu=4 * M_all * K_loc * 128 / us_sh / 1e6,
                   replicated_us=us_rep, replicated_gbs_per_gpu=b_rep / us_rep / 1e3, replicated_frac_hbm=b_rep / us_rep / 1e3 / peak,
                   replicated_tflops_per_gpu=4 * M * K * 128 / us_rep / 1e6,
                   loss_sharded=loss_sh, loss_replicated=loss_rep)
        rows.append(row)
        if rank == 0:
            print(f"G={world} K={K:8d} M/GPU={M:4d}: sharded K/G={K_loc:7d} x M_all={M_all:5d} {us_sh:8.1f} us/op "
                  f"{row['sharded_gbs_per_gpu']:6.0f} GB/s/GPU ({100 * row['sharded_frac_hbm']:4.1f}%) {row['sharded_tflops_per_gpu']:6.1f} TF/GPU | "
                  f"replicated {us_rep:8.1f} us/op ({100 * row['replicated_frac_hbm']:4.1f}%) | speed-up {us_rep / us_sh:4.2f}x | "
                  f"loss {loss_sh:.5f} / {loss_rep:.5f}", flush=True)
    del sharded, replicated
    torch.cuda.empty_cache()
if rank == 0 and args.out:
    os.makedirs(os.path.dirname(os.path.abspath(args.out)) or ".", exist_ok=True)
    json.dump(dict(hbm_peak_gbs=peak, exchange=fx.EXCHANGE, rows=rows), open(args.out, "w"), indent=1)
dist.destroy_process_group()
