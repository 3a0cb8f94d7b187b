"""World-size-2 tests of the host-side multi-GPU logic on the CPU (gloo): permutation broadcast,
shuffle all-to-all plan, rank-major key gather and shard ownership of the enqueue, fixed-shift partial
softmax statistics summed with reduce-scatter, and the single-collective `_parse_losses`.
The device kernels are replaced by their index-level definition (x[idx]) -- what is under test is the
routing and the bookkeeping, which are the same code on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, fn, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        out[rank] = fn(rank)
    finally:
        dist.destroy_process_group()


def _run(fn):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), fn, out), nprocs=WORLD, join=True)
    return [out[r] for r in range(WORLD)]


def _gather_cpu(x, idx):
    return x[idx]


# ---------------------------------------------------------------- shuffle-BN (moco.py:146-191)
def _shuffle_job(rank):
    from mscl_b200.recognizers import shuffle as shf
    from oracle import mscl_oracle as O
    n = 6
    torch.manual_seed(100 + rank)                 # ranks draw DIFFERENT permutations; rank 0's must win
    idx = shf.draw_permutation(n * WORLD, torch.device("cpu"))
    x_all = torch.arange(n * WORLD * 3, dtype=torch.float32).view(n * WORLD, 3)      # what all_gather would give
    x_local = x_all[rank * n:(rank + 1) * n].clone()
    plan = shf.ShufflePlan(idx, n, rank, WORLD)
    mine = shf.exchange(x_local, plan, _gather_cpu)
    want, unshuf = O.batch_shuffle(x_all, idx, rank, WORLD)
    # key features come back in the original order
    feats = mine * 2.0
    back = shf.exchange(feats, shf.ShufflePlan(unshuf, n, rank, WORLD), _gather_cpu)
    return dict(idx=idx.numpy(), ok=bool(torch.equal(mine, want)), back_ok=bool(torch.equal(back, x_local * 2.0)),
                sent=int(sum(plan.in_splits)))


def test_shuffle_plan_matches_gather_then_index():
    res = _run(_shuffle_job)
    np.testing.assert_array_equal(res[0]["idx"], res[1]["idx"])
    torch.manual_seed(100)
    np.testing.assert_array_equal(res[0]["idx"], torch.randperm(6 * WORLD).numpy())      # rank 0's draw
    assert all(r["ok"] and r["back_ok"] for r in res)
    assert all(r["sent"] == 6 for r in res)         # each rank ships exactly its n rows, not G*n


# ---------------------------------------------------------------- enqueue ownership (moco.py:423-440)
def _enqueue_job(rank):
    from mscl_b200.recognizers.moco import concat_all_gather
    from oracle import mscl_oracle as O
    K, n, C = 32, 4, 8
    K_local, begin = K // WORLD, rank * (K // WORLD)
    g = torch.Generator().manual_seed(5)
    queue_ref = torch.randn(C, K, generator=g)
    count_ref = torch.zeros(K, dtype=torch.long)
    shard = queue_ref[:, begin:begin + K_local].t().clone()          # key-major shard
    birth = torch.zeros(K_local, dtype=torch.int32)
    ptr = ptr_ref = 0
    n_enq = 0
    for step in range(6):                                             # wraps past K
        keys_local = torch.randn(n, C, generator=torch.Generator().manual_seed(10 * step + rank))
        keys = concat_all_gather(keys_local)                          # rank-major
        assert keys.shape[0] == n * WORLD
        both = torch.cat([torch.randn(n, C, generator=torch.Generator().manual_seed(10 * step + r)) for r in range(WORLD)])
        assert torch.equal(keys, both)
        # the enqueue kernel's index rule (csrc/enqueue.cu): slot = (ptr + i) % K, written iff inside this shard
        for i in range(keys.shape[0]):
            slot = (ptr + i) % K
            if begin <= slot < begin + K_local:
                shard[slot - begin] = keys[i]
                birth[slot - begin] = n_enq
        ptr = (ptr + keys.shape[0]) % K
        n_enq += 1
        ptr_ref = O.enqueue(queue_ref, count_ref, ptr_ref, keys)
    count = n_enq - birth.long()
    return dict(ok=bool(torch.equal(shard.t(), queue_ref[:, begin:begin + K_local])),
                count_ok=bool(torch.equal(count, count_ref[begin:begin + K_local])), ptr_ok=ptr == ptr_ref)


def test_sharded_enqueue_bookkeeping():
    assert all(r["ok"] and r["count_ok"] and r["ptr_ok"] for r in _run(_enqueue_job))


# ---------------------------------------------------------------- sharded InfoNCE statistics
def _lse_job(rank):
    """all_gather(qpack) -> partial (O, sum-exp, count) over the local shard with the FIXED shift ->
    reduce_scatter(SUM) -> finalize; must equal the replicated-queue loss (functional._InfoNCE)."""
    from oracle import mscl_oracle as O
    import torch.nn.functional as F
    M, C, K, T = 4, 16, 64, 0.07
    g = torch.Generator().manual_seed(3)
    q_all = F.normalize(torch.randn(M * WORLD, C, generator=g), dim=1)
    k_all = F.normalize(torch.randn(M * WORLD, C, generator=g), dim=1)
    queue = F.normalize(torch.randn(C, K, generator=g), dim=0)
    count = torch.randint(0, 2000, (K,), generator=g)
    w = O.decayed_weight(queue, count)
    q, k = q_all[rank * M:(rank + 1) * M], k_all[rank * M:(rank + 1) * M]
    pos = (q * k).sum(1) / T
    shift = q.norm(dim=1) * 1.0 / T
    pack = torch.cat([q, pos[:, None], shift[:, None]], dim=1)
    packs = [torch.empty_like(pack) for _ in range(WORLD)]
    dist.all_gather(packs, pack)
    pack_all = torch.cat(packs)
    K_local = K // WORLD
    w_loc = w[:, rank * K_local:(rank + 1) * K_local]
    s = pack_all[:, :C] @ w_loc / T
    p = torch.exp(s - pack_all[:, C + 1:C + 2])
    acc = torch.cat([p @ w_loc.t(), p.sum(1, keepdim=True), (s > pack_all[:, C:C + 1]).sum(1, keepdim=True).float()], dim=1)
    mine = torch.empty(M, acc.shape[1])
    dist.reduce_scatter(mine, list(acc.chunk(WORLD)), op=dist.ReduceOp.SUM)
    e0 = torch.exp(pos - shift)
    Z = e0 + mine[:, C]
    loss = (shift + torch.log(Z) - pos).mean()
    dq = ((e0 / Z - 1)[:, None] * k + mine[:, :C] / Z[:, None]) / T / M
    ql = q.clone().requires_grad_(True)
    logits = O.infonce_logits(ql, k, w, T)
    ref = O.cross_entropy_torch(logits, torch.zeros(M, dtype=torch.long))
    ref.backward()
    cnt_ref = (logits[:, 1:] > logits[:, :1]).sum(1).float()
    return dict(loss=float(loss), ref=float(ref), gerr=float((dq - ql.grad).norm() / ql.grad.norm()),
                cnt_ok=bool(torch.equal(mine[:, C + 1], cnt_ref)))


def test_sharded_partial_statistics_sum_to_the_replicated_loss():
    for r in _run(_lse_job):
        assert abs(r["loss"] - r["ref"]) < 1e-5 * abs(r["ref"]) and r["gerr"] < 1e-5 and r["cnt_ok"], r


# ---------------------------------------------------------------- _parse_losses (base.py:275-308)
def _parse_job(rank):
    from mscl_b200.recognizers.base_moco import BaseMoCoRecognizer
    losses = dict(top1_acc=torch.tensor(0.25 * (rank + 1)), loss_cls=torch.tensor([1.0 + rank, 3.0 + rank]),
                  loss_pos=torch.tensor(0.5), top5_acc=torch.tensor(1.0))
    loss, log_vars = BaseMoCoRecognizer._parse_losses(losses)
    return dict(loss=float(loss), log_vars=dict(log_vars))


def test_parse_losses_one_collective():
    res = _run(_parse_job)
    assert res[0]["loss"] == pytest.approx(2.5) and res[1]["loss"] == pytest.approx(3.5)      # local loss drives backward
    for r in res:                                                                            # log vars are rank means
        assert list(r["log_vars"]) == ["top1_acc", "loss_cls", "loss_pos", "top5_acc", "loss"]
        assert r["log_vars"]["top1_acc"] == pytest.approx(0.375)
        assert r["log_vars"]["loss_cls"] == pytest.approx(2.5)
        assert r["log_vars"]["loss"] == pytest.approx(3.0)


# ---------------------------------------------------------------- a whole recognizer step on two ranks
def _whole_step_job(rank):
    """MoCoV2.train_step on 2 ranks (different clips per rank): rank 0's permutation wins, the key batch is exchanged
    by the all-to-all plan, keys are gathered rank-major for ONE enqueue, iters advance by the GLOBAL batch, log
    variables are averaged over ranks -- against the oracle's in-process emulation of both ranks.  The four kernel
    entry points are replaced by oracle arithmetic as in tests/test_host_logic_cpu.py (test infrastructure)."""
    import mscl_b200
    from mscl_b200 import functional as fx
    from mscl_b200.recognizers.moco import MoCoV2, concat_all_gather
    from oracle import mscl_oracle as O
    from oracle.step import OracleMoCo
    import test_host_logic_cpu as H

    def fake_enqueue(self, keys):
        keys = concat_all_gather(keys.contiguous())                      # rank-major (moco.py:558-568)
        st = self._cpu_state
        st.ptr = O.enqueue(st.queue, st.count, st.ptr, keys.detach())
        self.batch_size = keys.shape[0]

    MoCoV2.contrast = H._fake_contrast
    MoCoV2._dequeue_and_enqueue = fake_enqueue
    fx.EmaTable = H._FakeEma
    fx.gather_rows = lambda x, idx: x[idx]
    cfg = dict(type="MoCoV2", backbone=dict(type="resnet_flow.r2d_18"), neck=dict(type="BaseMoCo"),
               moco_head=dict(type="MoCoHead", basename="", loss_cls=dict(type="CrossEntropyLoss_torch", ignore_index=-1)),
               im_key="imgs", dim_in=128, dim=128, K=64, m_base=0.99, max_iters=100, T=0.07, mlp=True, aux_info=[],
               aug=dict(type="IdentityAug"))
    torch.manual_seed(0)                                                  # same weights and queue on both ranks
    model = mscl_b200.build_model(cfg).train()
    orc = OracleMoCo(model)
    H._attach_cpu_state(model)
    n = 4
    ok, msgs = True, []
    for step in range(2):
        batches = [[torch.rand(n, 3, 8, 32, 32, generator=torch.Generator().manual_seed(1000 * step + 10 * r + v)) for v in range(2)]
                   for r in range(WORLD)]
        torch.manual_seed(200 + step)
        _, vars_ref = orc.train_step_ranks([b[0] for b in batches], [b[1] for b in batches])
        torch.manual_seed(200 + step + 7 * rank)                          # ranks draw DIFFERENT permutations; rank 0's must win
        if rank == 0:
            torch.manual_seed(200 + step)
        out = model.train_step(dict(imgs=batches[rank]), None)
        for k, v in out["log_vars"].items():
            if abs(v - vars_ref[k]) > 5e-5 * max(1.0, abs(vars_ref[k])):
                ok = False
                msgs.append((step, k, v, vars_ref[k]))
        st, ref = model._cpu_state, orc.branch.state
        if not (st.ptr == ref.ptr == (n * WORLD * (step + 1)) % 64 and model.iters == ref.iters == n * WORLD * (step + 1)
                and model.batch_size == n * WORLD and torch.equal(st.count, ref.count)
                and torch.allclose(st.queue, ref.queue, rtol=0, atol=1e-6)):
            ok = False
            msgs.append((step, "queue state", st.ptr, ref.ptr, model.iters, ref.iters))
        for pk, po in zip([p for m in (model.encoder_k, model.neck_k, model.mlp_k) for p in m.parameters()], orc.branch.k_params()):
            if not torch.equal(pk.detach(), po.detach()):
                ok = False
                msgs.append((step, "ema"))
                break
    return dict(ok=ok, msgs=msgs)


def test_whole_step_on_two_ranks_matches_the_oracle():
    for r in _run(_whole_step_job):
        assert r["ok"], r["msgs"]
