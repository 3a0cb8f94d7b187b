"""Parity of every CUDA kernel, called through the C-ABI, against the oracle and the golden
fixtures produced by the unmodified reference.  Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mscl_b200 import functional
    return functional


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ------------------------------------------------------------------ K5 enqueue (bit-exact)
def test_enqueue_golden(fx, golden_dir):
    g = _load(golden_dir, "enqueue_seq.npz")
    K = g["queue0"].shape[1]
    nq = fx.NegativeQueue(K)
    nq.load(torch.from_numpy(g["queue0"]), torch.zeros(K, dtype=torch.long), 0)
    keys = torch.from_numpy(g["keys"]).cuda()
    for s in range(keys.shape[0]):
        nq.enqueue(keys[s].contiguous())
        assert nq.ptr == int(g["ptrs"][s])
    q, c = nq.export()
    torch.cuda.synchronize()
    assert int(nq.qstate[0]) == int(g["ptrs"][-1]) and int(nq.qstate[1]) == keys.shape[0]
    np.testing.assert_array_equal(q.cpu().numpy(), g["queue"])
    np.testing.assert_array_equal(c.cpu().numpy(), g["count"])
    np.testing.assert_allclose(nq.weight().cpu().numpy(), g["weight"], rtol=2e-6, atol=0)


def test_enqueue_sharded_and_roundtrip(fx):
    from oracle import mscl_oracle as O
    K, B, G = 512, 32, 4
    gen = torch.Generator().manual_seed(3)
    q0 = F.normalize(torch.randn(128, K, generator=gen), dim=0)
    count0 = torch.randint(0, 50, (K,), generator=gen)
    shards = [fx.NegativeQueue(K, rank=r, world=G, shard=True) for r in range(G)]
    for s in shards:
        s.load(q0, count0, 7 * B)
    q_ref, c_ref, ptr = q0.clone(), count0.clone(), 7 * B
    for step in range(20):   # wraps past K
        keys = torch.randn(B, 128, generator=gen)
        ptr = O.enqueue(q_ref, c_ref, ptr, keys)
        kd = keys.cuda()
        for s in shards:
            s.enqueue(kd)
    qs, cs = zip(*[s.export() for s in shards])
    np.testing.assert_array_equal(torch.cat(qs, 1).cpu().numpy(), q_ref.numpy())
    np.testing.assert_array_equal(torch.cat(cs).cpu().numpy(), c_ref.numpy())
    assert all(s.ptr == ptr for s in shards)


def test_enqueue_rejects_bad_batch(fx):
    nq = fx.NegativeQueue(256)
    with pytest.raises(AssertionError):
        nq.enqueue(torch.zeros(48, 128, device="cuda"))


# ------------------------------------------------------------------ K4 EMA (bit-exact)
def test_ema_golden(fx, golden_dir):
    g = _load(golden_dir, "ema.npz")
    names = [str(n) for n in g["names"]]
    ks = [torch.from_numpy(g[f"k0/{n}"]).cuda() for n in names]
    qs = [torch.from_numpy(g[f"q/{n}"]).cuda() for n in names]
    tab = fx.EmaTable(ks, qs)
    for step in range(3):
        tab.update(float(g["m"][step]))
        for n, k in zip(names, ks):
            np.testing.assert_array_equal(k.cpu().numpy(), g[f"k{step + 1}/{n}"])


def test_ema_large_and_ragged(fx):
    from oracle import mscl_oracle as O
    gen = torch.Generator().manual_seed(0)
    sizes = [1, 3, 64, 1000, 16384, 16385, 3 * 16384 + 7, 2_000_003]
    ks = [torch.randn(n, generator=gen) for n in sizes]
    qs = [torch.randn(n, generator=gen) for n in sizes]
    m = O.momentum(12345, 100000, 0.994)
    ref = O.ema_update(ks, qs, m)
    kd = [k.cuda() for k in ks]
    # an unaligned view exercises the scalar path
    big = torch.randn(4099, generator=gen)
    kd.append(big.cuda()[3:])
    qs.append(torch.randn(4096, generator=gen))
    ref.append(O.ema_update([big[3:]], [qs[-1]], m)[0])
    tab = fx.EmaTable(kd, [q.cuda() for q in qs])
    tab.update(m)
    for a, b in zip(kd, ref):
        np.testing.assert_array_equal(a.cpu().numpy(), b.numpy())


# ------------------------------------------------------------------ K3 FRA
def test_fra_golden(fx, golden_dir):
    g = _load(golden_dir, "fra.npz")
    tab = fx.fra_table()
    for i in range(2):
        x = torch.from_numpy(g[f"in{i}"]).cuda()           # (T,H,W,2)
        cid = torch.tensor([int(g[f"cid{i}"])], dtype=torch.int32, device="cuda")
        ref = g[f"out{i}"]                                  # (2T,H,W,2)
        for layout in ("interleaved", "planar"):
            inp = x[None].contiguous() if layout == "interleaved" else x.permute(3, 0, 1, 2)[None].contiguous()
            out = fx.fra(inp, cid, tab, layout)             # (1,2,2T,H,W)
            got = out[0].permute(1, 2, 3, 0).cpu().numpy()
            np.testing.assert_allclose(got, ref, rtol=3e-6, atol=3e-7)


@pytest.mark.parametrize("one_pass,hw", [(True, (112, 112)), (False, (112, 112)), (True, (36, 52)), (True, (224, 224))])
def test_fra_oracle_bit_exact_full_size(fx, one_pass, hw):
    """one_pass=True: the cluster kernel (frame staged in distributed shared memory); False: max pre-pass + apply."""
    from oracle import inputs, mscl_oracle as O
    N, T, (H, W) = 3, 8, hw
    rs = np.random.RandomState(0)
    cids = rs.randint(0, 8, size=N)
    clips = [inputs.flow_clip(seed=10 + n, T=T, H=H, W=W) for n in range(N)]
    ref = np.stack([np.stack(O.fra(c, int(k))) for c, k in zip(clips, cids)])      # (N,2T,H,W,2)
    x = torch.from_numpy(np.stack([np.stack(c) for c in clips])).cuda()            # (N,T,H,W,2)
    out = fx.fra(x, torch.from_numpy(cids.astype(np.int32)).cuda(), fx.fra_table(), "interleaved", one_pass=one_pass)
    got = out.permute(0, 2, 3, 4, 1).cpu().numpy()
    np.testing.assert_array_equal(got, ref)   # same float32 operation sequence -> same bits
    planar = x.permute(0, 4, 1, 2, 3).contiguous()
    out_p = fx.fra(planar, torch.from_numpy(cids.astype(np.int32)).cuda(), fx.fra_table(), "planar", one_pass=one_pass)
    np.testing.assert_array_equal(out_p.cpu().numpy(), out.cpu().numpy())
    rot = fx.fra_rotate(out[:, :, :T].contiguous(), torch.from_numpy(cids.astype(np.int32)).cuda(), fx.fra_table())
    np.testing.assert_allclose(rot.permute(0, 2, 3, 4, 1).cpu().numpy(), ref[:, T:], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ K2 LMCL
@pytest.mark.parametrize("shape", [(2, 128, 8, 7, 7), (3, 5, 3, 7, 7), (2, 128, 4, 28, 28), (1, 7, 3, 2, 2), (2, 3, 5, 1, 3),
                                   (2, 16, 4, 10, 10), (1, 130, 1, 9, 14)])
def test_hw_mean_matches_torch(fx, shape):
    """AdaptiveAvgPool3d((None,1,1)) (local_cl_head.py:61-62): both kernels (warp-per-row, and the staged form
    used for planes under 128 pixels), forward and backward, ragged row counts."""
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).cuda().requires_grad_(True)
    w = torch.randn(shape[:3], generator=g).cuda()
    out = fx.hw_mean(x)
    ref = x.detach().double().mean(dim=(-2, -1))
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-6)
    (out * w).sum().backward()
    want = (w / (shape[-1] * shape[-2]))[..., None, None].expand(shape)
    np.testing.assert_allclose(x.grad.cpu().numpy(), want.cpu().numpy(), rtol=1e-6, atol=0)


@pytest.mark.parametrize("shape", [(2, 128, 4, 28, 28), (3, 128, 8, 7, 7), (2, 32, 3, 5, 6), (1, 64, 1, 9, 14), (2, 256, 2, 1, 3),
                                   (2, 48, 2, 4, 4)])
def test_hw_mean_channels_last(fx, shape):
    """The NDHWC kernels: a channels_last_3d feature map is pooled where it lies (no layout copy) and its gradient comes
    back channels_last_3d; C not a multiple of 32 (last case) takes the row-major kernels, values unchanged."""
    g = torch.Generator().manual_seed(sum(shape) + 3)
    cl = torch.channels_last_3d
    x = torch.randn(shape, generator=g).cuda().contiguous(memory_format=cl).requires_grad_(True)
    w = torch.randn(shape[:3], generator=g).cuda()
    out = fx.hw_mean(x)
    assert out.is_contiguous() and tuple(out.shape) == shape[:3]
    ref = x.detach().double().mean(dim=(-2, -1))
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-6)
    (out * w).sum().backward()
    if shape[1] % 32 == 0 and shape[2] * shape[3] * shape[4] > 1:
        assert x.grad.is_contiguous(memory_format=cl)
    want = (w / (shape[-1] * shape[-2]))[..., None, None].expand(shape)
    np.testing.assert_allclose(x.grad.cpu().numpy(), want.cpu().numpy(), rtol=1e-6, atol=0)


def _oracle_lmcl(q_map, qf_map, qaf_map, T, t):
    from oracle import mscl_oracle as O
    leaves = [x.clone().requires_grad_(True) for x in (q_map, qf_map, qaf_map)]
    out = O.lmcl(*leaves, T, t)
    out["loss_pos"].backward()
    return out, [l.grad for l in leaves]


@pytest.mark.parametrize("N,t,hw_rgb,hw_flow", [(4, 4, 6, 3), (8, 8, 28, 7), (16, 16, 14, 7)])
def test_lmcl_vs_oracle(fx, N, t, hw_rgb, hw_flow):
    from oracle import inputs
    inp = inputs.head_inputs(seed=2, N=N, K=64, t=t, hw_rgb=hw_rgb, hw_flow=hw_flow)
    ref, gref = _oracle_lmcl(inp["q_map"], inp["qf_map"], inp["qaf_map"], 0.07, t)
    maps = [inp[k].cuda().requires_grad_(True) for k in ("q_map", "qf_map", "qaf_map")]
    xq = fx.hw_mean(maps[0])
    xf = fx.hw_mean(torch.cat((maps[1], maps[2]), dim=2))
    out = fx.lmcl(xq, xf, 0.07)
    out[0].backward()
    o = out.detach().cpu()
    assert abs(float(o[0]) - float(ref["loss_pos"])) <= 1e-4 * abs(float(ref["loss_pos"]))   # 1e-3 contract, fp32 path
    assert float(o[1]) == pytest.approx(float(ref["top1_acc_pos"]), abs=1e-6)
    assert float(o[2]) == pytest.approx(float(ref["top5_acc_pos"]), abs=1e-6)
    for m, g in zip(maps, gref):
        assert _rel(m.grad.cpu(), g) < 1e-4


# ------------------------------------------------------------------ K1 InfoNCE
def _oracle_infonce(q, kpos, queue_ck, count, T, rpg):
    from oracle import mscl_oracle as O
    ql = q.clone().requires_grad_(True)
    w = O.decayed_weight(queue_ck, count)
    logits = O.infonce_logits(ql, kpos, w, T)
    labels = torch.zeros(rpg, dtype=torch.long)
    res = []
    for g0 in range(0, q.shape[0], rpg):
        lg = logits[g0:g0 + rpg]
        acc = O.top_k_accuracy(lg.detach().numpy(), labels.numpy(), (1, 5))
        res.append((O.cross_entropy_torch(lg, labels), acc[0], acc[1]))
    sum(r[0] for r in res).backward()
    return res, ql.grad, logits.detach()


def _make_case(seed, M, K, b_all):
    from oracle import inputs
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(M, 128, generator=g)
    noise = torch.linspace(0.15, 1.6, M).unsqueeze(1)
    q = F.normalize(z + noise * torch.randn(M, 128, generator=g), dim=1)
    kpos = F.normalize(z + noise * torch.randn(M, 128, generator=g), dim=1)
    queue = F.normalize(torch.randn(128, K, generator=g), dim=0)
    count = inputs.steady_state_count(K, b_all, (5 * b_all) % K)
    return q, kpos, queue, count


CASES = [  # M, K, rows_per_group, b_all
    (4, 256, 4, 4), (8, 4096, 8, 8), (24, 1000, 8, 8), (96, 65536, 32, 128), (192, 16384, 64, 64), (130, 8192, 65, 64),
]


@pytest.mark.parametrize("impl", ["simt", "tc", "fused", "fused_pass"])
@pytest.mark.parametrize("M,K,rpg,b_all", CASES)
def test_infonce_vs_oracle(fx, impl, M, K, rpg, b_all):
    q, kpos, queue, count = _make_case(M + K, M, K, b_all)
    ref, gref, logits = _oracle_infonce(q, kpos, queue, count, 0.07, rpg)
    nq = fx.NegativeQueue(K)
    nq.load(queue, count, (5 * b_all) % K)
    qd = q.cuda().requires_grad_(True)
    out, rows = fx.infonce(qd, kpos.cuda(), nq, rpg, 0.07, impl=impl)
    out[:, 0].sum().backward()
    o = out.detach().cpu()
    # margin between the positive and its nearest-ranked negative decides whether a top-k flag may flip
    neg = logits[:, 1:]
    pos = logits[:, :1]
    # contract: 1e-3 relative (tf32 operands, fp32 accumulate); the CUDA-core twin only sees tf32-rounded q
    tol = 3e-4 if impl == "simt" else 1e-3
    for gi, (loss, t1, t5) in enumerate(ref):
        assert abs(float(o[gi, 0]) - float(loss)) <= tol * abs(float(loss)), (gi, float(o[gi, 0]), float(loss))
    cnt_ref = (neg > pos).sum(1).float()
    cnt = rows[M:].cpu()
    close_call = ((neg - pos).abs() < 0.02).sum(1)       # negatives within tf32 noise of the positive
    assert bool(((cnt - cnt_ref).abs() <= close_call).all()), (cnt, cnt_ref)
    for gi, (loss, t1, t5) in enumerate(ref):
        sl = slice(gi * rpg, (gi + 1) * rpg)
        if int(close_call[sl].sum()) == 0:
            assert float(o[gi, 1]) == pytest.approx(t1, abs=1e-6) and float(o[gi, 2]) == pytest.approx(t5, abs=1e-6)
    assert _rel(qd.grad.cpu(), gref) < tol, _rel(qd.grad.cpu(), gref)


@pytest.mark.parametrize("impl", ["tc", "fused", "fused_pass"])
def test_infonce_tc_matches_simt_no_grad(fx, impl):
    M, K = 64, 32768
    q, kpos, queue, count = _make_case(9, M, K, 64)
    nq = fx.NegativeQueue(K)
    nq.load(queue, count, 0)
    with torch.no_grad():
        a, ra = fx.infonce(q.cuda(), kpos.cuda(), nq, 32, 0.07, impl=impl)
        b, rb = fx.infonce(q.cuda(), kpos.cuda(), nq, 32, 0.07, impl="simt")
    assert _rel(a[:, 0].cpu(), b[:, 0].cpu()) < 1e-3
    assert float((ra[M:] - rb[M:]).abs().max()) <= 2            # hit counts: tf32 tensor-core vs fp32 products of the same operands


@pytest.mark.parametrize("impl", ["tc", "fused"])
@pytest.mark.parametrize("K", [262144, 1048576])
def test_infonce_large_queue_vs_oracle(fx, impl, K):
    """The far end of BASELINE configs[2] (queue sweep to 1 Mi negatives) against the CPU oracle: loss, hit counts, dq."""
    M, rpg = 64, 32
    q, kpos, queue, count = _make_case(K // 1024, M, K, 256)
    ref, gref, logits = _oracle_infonce(q, kpos, queue, count, 0.07, rpg)
    nq = fx.NegativeQueue(K)
    nq.load(queue, count, (5 * 256) % K)
    qd = q.cuda().requires_grad_(True)
    out, rows = fx.infonce(qd, kpos.cuda(), nq, rpg, 0.07, impl=impl)
    out[:, 0].sum().backward()
    for gi, (loss, _, _) in enumerate(ref):
        assert abs(float(out[gi, 0]) - float(loss)) <= 1e-3 * abs(float(loss))
    neg, pos = logits[:, 1:], logits[:, :1]
    close_call = ((neg - pos).abs() < 0.02).sum(1)
    assert bool(((rows[M:].cpu() - (neg > pos).sum(1).float()).abs() <= close_call).all())
    assert _rel(qd.grad.cpu(), gref) < 1e-3


def test_infonce_multi_jobs_match_single_launches_and_oracle(fx):
    """mscl_infonce_fused_multi: independent terms over different queues in ONE launch (what MSCLWithAug.objective does
    with its two pre-enqueue passes) -- every job's loss / hit counts / dq against the oracle and against the same job
    launched alone; jobs of different sizes, one with a positive key present in its queue."""
    specs = [(96, 65536, 32, 128, 0.07), (32, 65536, 32, 128, 0.07), (24, 1000, 8, 8, 0.2)]
    cases, jobs, qds = [], [], []
    for i, (M, K, rpg, b_all, T) in enumerate(specs):
        q, kpos, queue, count = _make_case(100 + i, M, K, b_all)
        dup = None
        if i == 1:      # rows 0..M-1 find their own positive at slots 5*b_all .. (the rf term of MSCLWithAug)
            ptr = 5 * b_all
            queue[:, ptr:ptr + M] = kpos.t()
            count = count + 1
            count[ptr:ptr + M] = 1
            dup = (torch.arange(M, dtype=torch.int32) + ptr).cuda()
        nq = fx.NegativeQueue(K)
        nq.load(queue, count, 0)
        qd = q.cuda().requires_grad_(True)
        cases.append((q, kpos, queue, count, T, rpg, M))
        qds.append(qd)
        jobs.append(dict(q=qd, kpos=kpos.cuda(), nq=nq, rows_per_group=rpg, T=T, dup_slot=dup, dup_age=1))
    outs = fx.infonce_multi(jobs)
    sum(o[:, 0].sum() for o, _ in outs).backward()
    for (q, kpos, queue, count, T, rpg, M), job, qd, (out, rows) in zip(cases, jobs, qds, outs):
        ref, gref, logits = _oracle_infonce(q, kpos, queue, count, T, rpg)
        for gi, (loss, _, _) in enumerate(ref):
            assert abs(float(out[gi, 0]) - float(loss)) <= 1e-3 * abs(float(loss))
        neg, pos = logits[:, 1:].clone(), logits[:, :1]
        if job["dup_slot"] is not None:
            neg[torch.arange(M), job["dup_slot"].cpu().long()] = 1e9
        close_call = ((neg - pos).abs() < 0.02 * (0.07 / T) + 1e-3).sum(1)
        cnt_ref = (logits[:, 1:] > logits[:, :1]).sum(1).float()
        assert bool(((rows[M:].cpu() - cnt_ref).abs() <= close_call).all())
        assert _rel(qd.grad.cpu(), gref) < 1e-3
        # the same job alone
        q1 = q.cuda().requires_grad_(True)
        o1, r1 = fx.infonce(q1, job["kpos"], job["nq"], rpg, T, impl="fused", dup_slot=job["dup_slot"], dup_age=1)
        o1[:, 0].sum().backward()
        assert _rel(o1[:, 0], out[:, 0]) < 1e-6 and torch.equal(r1[M:], rows[M:]) and _rel(q1.grad, qd.grad) < 1e-5


@pytest.mark.parametrize("n,K,ptr_steps,with_other_job", [(32, 65536, 5, True), (32, 65536, 2047, False), (8, 4096, 3, False),
                                                          (64, 16384, 1, True), (24, 960, 0, False), (128, 8192, 7, True)])
def test_infonce_epoch_split_vs_oracle(fx, n, K, ptr_steps, with_other_job):
    """mscl_infonce_fused_multi_x: ONE pass over a queue for rows that read it as it was before its last enqueue (the
    base-flow call's own term, moco.py:481-498) and rows that read it after (mscl.py:239-277: own_aug, rf -- whose positives
    ARE the new keys -- and rf_aug).  Checked against the oracle evaluated on the two queue states, and against the
    three-pass schedule (pass, enqueue, pass) on the device."""
    from oracle import mscl_oracle as O
    T = 0.07
    q_all, kpos_all, queue, count = _make_case(n + K + ptr_steps, 4 * n, K, n)
    ptr = (ptr_steps * n) % K
    g = torch.Generator().manual_seed(n)
    new_keys = F.normalize(torch.randn(n, 128, generator=g), dim=1)
    kpos_all[2 * n:3 * n] = new_keys                      # the rf group's positives are the keys about to be enqueued
    q_all[2 * n:3 * n] = F.normalize(new_keys + torch.linspace(0.05, 1.2, n).unsqueeze(1) * torch.randn(n, 128, generator=g), dim=1)
    # oracle: pre rows on the queue as it is, post rows on the queue after the enqueue
    ref_pre, g_pre, lg_pre = _oracle_infonce(q_all[:n], kpos_all[:n], queue, count, T, n)
    queue2, count2 = queue.clone(), count.clone()
    assert O.enqueue(queue2, count2, ptr, new_keys) == (ptr + n) % K
    ref_post, g_post, lg_post = _oracle_infonce(q_all[n:], kpos_all[n:], queue2, count2, T, n)
    ref, gref = ref_pre + ref_post, torch.cat([g_pre, g_post])
    dup = torch.full((4 * n,), -1, dtype=torch.int32)
    dup[2 * n:3 * n] = torch.arange(n, dtype=torch.int32) + ptr
    nq = fx.NegativeQueue(K)
    nq.load(queue, count, ptr)
    overwritten = nq.enqueue(new_keys.cuda(), save=True)
    assert overwritten[2] == ptr and torch.equal(overwritten[1].cpu().long(), int(count.max()) - count[ptr:ptr + n])
    qd = q_all.cuda().requires_grad_(True)
    jobs = [dict(q=qd, kpos=kpos_all.cuda(), nq=nq, rows_per_group=n, T=T, dup_slot=dup.cuda(), dup_age=1,
                 overwritten=overwritten, row_split=n)]
    if with_other_job:       # as in the step: an ordinary job over another queue in the same launch, listed first
        q2, k2, queue_b, count_b = _make_case(7, 96, 65536, 32)
        nq_b = fx.NegativeQueue(65536)
        nq_b.load(queue_b, count_b, 0)
        q2d = q2.cuda().requires_grad_(True)
        jobs.insert(0, dict(q=q2d, kpos=k2.cuda(), nq=nq_b, rows_per_group=32, T=T))
    outs = fx.infonce_multi(jobs)
    sum(o[:, 0].sum() for o, _ in outs).backward()
    out, rows = outs[-1]
    for gi, (loss, _, _) in enumerate(ref):
        assert abs(float(out[gi, 0]) - float(loss)) <= 1e-3 * abs(float(loss)), (gi, float(out[gi, 0]), float(loss))
    logits = torch.cat([lg_pre, lg_post])
    neg, pos = logits[:, 1:].clone(), logits[:, :1]
    cnt_ref = (neg > pos).sum(1).float()
    neg[torch.arange(2 * n, 3 * n), dup[2 * n:3 * n].long()] = 1e9      # ranked exactly by the epilogue, never a close call
    close_call = ((neg - pos).abs() < 0.02).sum(1)
    assert bool(((rows[4 * n:].cpu() - cnt_ref).abs() <= close_call).all()), (rows[4 * n:].cpu(), cnt_ref)
    assert _rel(qd.grad.cpu(), gref) < 1e-3, _rel(qd.grad.cpu(), gref)
    if with_other_job:
        ref_b, g_b, _ = _oracle_infonce(q2, k2, queue_b, count_b, T, 32)
        for gi, (loss, _, _) in enumerate(ref_b):
            assert abs(float(outs[0][0][gi, 0]) - float(loss)) <= 1e-3 * abs(float(loss))
        assert _rel(q2d.grad.cpu(), g_b) < 1e-3
    # the three-pass schedule on the device: pass over the pre rows, enqueue, pass over the post rows
    q3 = q_all.cuda().requires_grad_(True)
    kd = kpos_all.cuda()
    nq3 = fx.NegativeQueue(K)
    nq3.load(queue, count, ptr)
    o_pre, r_pre = fx.infonce(q3[:n], kd[:n], nq3, n, T)
    nq3.enqueue(new_keys.cuda())
    assert torch.equal(nq3.queue, nq.queue) and torch.equal(nq3.birth, nq.birth)
    o_post, r_post = fx.infonce(q3[n:], kd[n:], nq3, n, T, dup_slot=dup[n:].cuda(), dup_age=1)
    (o_pre[:, 0].sum() + o_post[:, 0].sum()).backward()
    assert _rel(torch.cat([o_pre[:, 0], o_post[:, 0]]), out[:, 0]) < 2e-4
    assert float((torch.cat([r_pre[n:], r_post[3 * n:]]) - rows[4 * n:]).abs().max()) <= 2
    assert _rel(q3.grad, qd.grad) < 5e-4


def test_infonce_fused_workspace_stays_zero_and_repeats(fx):
    """mscl_infonce_fused accumulates the row statistics into a zero workspace and must leave it zero (accumulator AND
    CTA counter), so back-to-back calls agree to rounding (the float adds are unordered) and never see stale sums."""
    M, K = 96, 65536
    q, kpos, queue, count = _make_case(3, M, K, 128)
    nq = fx.NegativeQueue(K)
    nq.load(queue, count, 0)
    outs = []
    for _ in range(4):
        qd = q.cuda().requires_grad_(True)
        out, rows = fx.infonce(qd, kpos.cuda(), nq, 32, 0.07, impl="fused")
        out[:, 0].sum().backward()
        outs.append((out.detach().clone(), rows.clone(), qd.grad.clone()))
    ws = fx._fused_workspace(q.cuda().device, M)
    torch.cuda.synchronize()
    assert int(torch.count_nonzero(ws)) == 0
    for o, r, g in outs[1:]:
        assert _rel(o[:, 0], outs[0][0][:, 0]) < 1e-6
        assert torch.equal(r[M:], outs[0][1][M:])                # hit counts are integers: exact whatever the add order
        assert _rel(g, outs[0][2]) < 1e-6                        # the slabs are summed in a fixed order


def test_infonce_fresh_queue_and_after_enqueue(fx):
    """count == 0 everywhere (moco.py:396) and the snapshot-after-enqueue case (App. A.2)."""
    from oracle import mscl_oracle as O
    M, K = 16, 1024
    q, kpos, queue, _ = _make_case(4, M, K, 16)
    count = torch.zeros(K, dtype=torch.long)
    nq = fx.NegativeQueue(K)
    nq.load(queue, count, 0)
    keys = F.normalize(torch.randn(16, 128, generator=torch.Generator().manual_seed(1)), dim=1)
    for step in range(3):
        ref, gref, _ = _oracle_infonce(q, kpos, queue, count, 0.07, M)
        qd = q.cuda().requires_grad_(True)
        out, _ = fx.infonce(qd, kpos.cuda(), nq, M, 0.07)
        out[0, 0].backward()
        assert abs(float(out[0, 0]) - float(ref[0][0])) <= 1e-3 * abs(float(ref[0][0]))
        assert _rel(qd.grad.cpu(), gref) < 1e-3
        O.enqueue(queue, count, nq.ptr, keys)
        nq.enqueue(keys.cuda())


@pytest.mark.parametrize("impl", ["simt", "tc", "fused", "fused_pass"])
def test_infonce_positive_key_present_in_queue(fx, impl):
    """The rf term of MSCLWithAug reads the flow queue right after k_flow was enqueued (mscl.py:239-248):
    every row finds its own positive key among the negatives, scored pos*0.99999.  The reference ranks it
    above the positive iff pos < 0; the kernel must reproduce that exactly (dup_slot)."""
    M, K, B = 32, 2048, 32
    q, kpos, queue, count = _make_case(21, M, K, B)
    kpos[::3] = -kpos[::3]                       # a third of the rows get a negative positive-logit
    ptr = 5 * B
    queue[:, ptr:ptr + M] = kpos.t()
    count = count + 1
    count[ptr:ptr + M] = 1
    ref, gref, logits = _oracle_infonce(q, kpos, queue, count, 0.07, M)
    cnt_ref = (logits[:, 1:] > logits[:, :1]).sum(1).float()
    assert int((logits[:, 1 + ptr:1 + ptr + M].diagonal() > logits[:, 0]).sum()) == len(range(0, M, 3))
    nq = fx.NegativeQueue(K)
    nq.load(queue, count, ptr + M)
    slots = (torch.arange(M, dtype=torch.int32) + ptr).cuda()
    qd = q.cuda().requires_grad_(True)
    out, rows = fx.infonce(qd, kpos.cuda(), nq, M, 0.07, impl=impl, dup_slot=slots, dup_age=1)
    out[0, 0].backward()
    neg, pos = logits[:, 1:].clone(), logits[:, :1]
    neg[torch.arange(M), ptr + torch.arange(M)] = 1e9          # the duplicate itself is never a close call
    close_call = ((neg - pos).abs() < 0.02).sum(1)
    assert bool(((rows[M:].cpu() - cnt_ref).abs() <= close_call).all()), (rows[M:].cpu(), cnt_ref)
    assert int(close_call.sum()) < M                            # the test is not vacuous
    assert abs(float(out[0, 0]) - float(ref[0][0])) <= 1e-3 * abs(float(ref[0][0]))
    assert _rel(qd.grad.cpu(), gref) < 1e-3


# ------------------------------------------------------------------ K6 gather
def test_gather_rows(fx):
    x = torch.randn(64, 3, 8, 28, 28, device="cuda")
    idx = torch.randperm(64)[:16].cuda()
    np.testing.assert_array_equal(fx.gather_rows(x, idx).cpu().numpy(), x[idx].cpu().numpy())


# ------------------------------------------------------------------ K7 trilinear up-sampling (TPN neck)
@pytest.mark.parametrize("shape,size", [((2, 8, 2, 14, 14), (4, 28, 28)), ((2, 5, 1, 7, 7), (2, 14, 14)), ((1, 3, 3, 5, 6), (7, 9, 11)),
                                        ((2, 4, 4, 6, 6), (4, 6, 6)), ((1, 2, 2, 3, 3), (5, 8, 12))])
def test_upsample_trilinear_matches_torch(fx, shape, size):
    """necks/sepc.py:126-130: F.interpolate(mode="trilinear", align_corners=False); forward and the gather backward."""
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).cuda().requires_grad_(True)
    w = torch.randn(shape[:2] + size, generator=g).cuda()
    y = fx.upsample_trilinear(x, size)
    xr = x.detach().clone().requires_grad_(True)
    yr = F.interpolate(xr, size=size, mode="trilinear")
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    (y * w).sum().backward()
    (yr * w).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), xr.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    xc = x.detach().cpu().requires_grad_(True)       # and against the host op the oracle uses
    yc = F.interpolate(xc, size=size, mode="trilinear")
    np.testing.assert_allclose(y.detach().cpu().numpy(), yc.detach().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shape,size", [((2, 128, 2, 14, 14), (4, 28, 28)), ((3, 128, 1, 7, 7), (2, 14, 14)), ((1, 8, 3, 5, 6), (7, 9, 11)),
                                        ((2, 4, 4, 6, 6), (4, 6, 6)), ((2, 12, 2, 3, 3), (5, 8, 12)),
                                        # C % 16 == 0: the one-pass backward, at non-integer scales and with an unscaled axis
                                        ((1, 16, 3, 5, 6), (7, 9, 11)), ((2, 32, 4, 6, 6), (4, 6, 6)), ((2, 48, 2, 3, 3), (5, 8, 12))])
def test_upsample_trilinear_channels_last(fx, shape, size):
    """The NDHWC kernels: a channels_last_3d input gives a channels_last_3d output and input gradient, same values as
    F.interpolate; the incoming gradient may arrive in either layout."""
    g = torch.Generator().manual_seed(sum(shape) + 1)
    cl = torch.channels_last_3d
    x = torch.randn(shape, generator=g).cuda().contiguous(memory_format=cl).requires_grad_(True)
    w = torch.randn(shape[:2] + size, generator=g).cuda()
    y = fx.upsample_trilinear(x, size)
    assert y.is_contiguous(memory_format=cl) and tuple(y.shape) == shape[:2] + size
    xr = x.detach().clone().contiguous().requires_grad_(True)
    yr = F.interpolate(xr, size=size, mode="trilinear")
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    (yr * w).sum().backward()
    for wgt in (w, w.contiguous(memory_format=cl)):          # row-major and channels-last incoming gradients
        x.grad = None
        (fx.upsample_trilinear(x, size) * wgt).sum().backward()
        assert x.grad.is_contiguous(memory_format=cl)
        np.testing.assert_allclose(x.grad.cpu().numpy(), xr.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    # C not a multiple of 4 falls back to the row-major kernels (values unchanged)
    x3 = torch.randn(2, 6, 2, 4, 4, generator=g).cuda().contiguous(memory_format=cl)
    np.testing.assert_allclose(fx.upsample_trilinear(x3, (4, 8, 8)).cpu().numpy(),
                               F.interpolate(x3.contiguous(), size=(4, 8, 8), mode="trilinear").cpu().numpy(), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ K8 / K9 augmentation front-end
def _flowvis_close(got, ref):
    """Every op of the visualiser is exactly rounded except atan2 (device vs host libm differ in the last ulp), which can
    move a pixel across a floor(255 * col) boundary: values are k/255, so a mismatch is exactly one level, and rare."""
    d = np.abs(got - ref)
    assert d.max() <= 1.0 / 255 + 1e-7, d.max()
    assert (d > 0).mean() <= 5e-3, (d > 0).mean()


def test_flow_visualize_golden_and_oracle(fx, golden_dir):
    from oracle import mscl_oracle as O
    g = _load(golden_dir, "flowvis.npz")
    got = fx.flow_visualize(torch.from_numpy(g["flows"]).cuda())
    _flowvis_close(got.cpu().numpy(), g["out"])
    gen = torch.Generator().manual_seed(3)
    flows = torch.randn(4, 2, 16, 112, 112, generator=gen) * torch.tensor([0.2, 0.6, 1.0, 3.0]).view(4, 1, 1, 1, 1)
    ref = O.flow_visualize(flows)
    flip = torch.tensor([1, 0, 1, 0], dtype=torch.uint8)
    got = fx.flow_visualize(flows.cuda(), flip.cuda())
    want = torch.where(flip.bool().view(-1, 1, 1, 1, 1), torch.flip(ref, [-1]), ref)    # ssl_aug_v2.py:109-117
    _flowvis_close(got.cpu().numpy(), want.numpy())
    norm = torch.tensor([0.485, 0.456, 0.406, 0.229, 0.224, 0.225])
    got_n = fx.flow_visualize(flows.cuda(), None, norm.cuda())
    want_n = (ref - norm[:3].view(1, 3, 1, 1, 1)) / norm[3:].view(1, 3, 1, 1, 1)
    d = np.abs(got_n.cpu().numpy() - want_n.numpy())
    assert d.max() <= (1.0 / 255) / 0.224 + 1e-5 and (d > 1e-6).mean() <= 5e-3


@pytest.mark.parametrize("per_frame", [False, True])
@pytest.mark.parametrize("shape,crop", [((6, 3, 4, 32, 48), 112), ((3, 3, 8, 112, 112), 112), ((6, 3, 2, 24, 36), 64),
                                        ((2, 3, 2, 130, 132), 112)])
def test_color_pipeline_matches_torch_ops(fx, shape, crop, per_frame):
    """K9 against the same pipeline written as PyTorch ops (oracle/aug_oracle.py): every
    combination of jitter / grayscale / blur / flip decisions, given parameters.  crop 112 -> 11 taps (frames up to 128
    wide take the in-place shared-memory kernel, wider ones the generic one), crop 64 -> 7 taps (generic kernel).
    per_frame: the jitter factors differ between the frames of a clip ('batch' sync level, ssl_aug.py:56-60) and the
    contrast step uses each frame's own mean luminance; otherwise one set per clip ('params' level, :62-66)."""
    from mscl_b200.common.ssl_aug import SyncMoCoAugmentV5
    from oracle import aug_oracle as A
    aug = SyncMoCoAugmentV5(crop_size=crop, sync_level=("batch", "batch"), t=(8, 8), flow_suffix="flow_imgs")
    mean, std = aug.mean.view(-1), aug.std.view(-1)
    n = shape[0]
    gen = torch.Generator().manual_seed(n)
    x = torch.rand(shape, generator=gen)
    torch.manual_seed(5)
    frames = shape[2] if per_frame else 1
    prm = aug._color_params(n, torch.device("cpu"), frames)
    assert prm["brightness"].numel() == n * frames
    combos = [(1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 0, 0)]
    rep = lambda j: torch.tensor([combos[i % 6][j] for i in range(n)], dtype=torch.bool).repeat_interleave(frames)
    prm["jit"], prm["gray"], prm["blur"] = rep(0), rep(1), rep(2)
    flip = torch.tensor([i % 2 == 0 for i in range(n)])
    want = A.normalize(A.color_pipeline(A.flip(x, flip), prm, aug.blur_radius), mean, std)
    dev = torch.device("cuda")
    prm_d = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in prm.items()}
    norm = torch.cat([aug.mean.view(-1), aug.std.view(-1)]).to(dev)
    flip_u = flip.repeat_interleave(frames).to(dev)
    got = fx.color_pipeline(x.to(dev), aug._pack_params(prm_d, flip_u, False), prm_d["taps"].contiguous(), norm)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-5, atol=2e-5)
    if per_frame and shape[2] > 1:      # the frames of a jittered clip really got different factors
        assert prm["brightness"].view(n, frames).std(dim=1).min() > 0
    weak = fx.color_pipeline(x.to(dev), aug._pack_params(prm_d, flip_u, True), prm_d["taps"].contiguous(), norm)
    np.testing.assert_allclose(weak.cpu().numpy(), A.normalize(A.flip(x, flip), mean, std).numpy(), rtol=1e-6, atol=1e-6)


def test_augmentation_module_on_device(fx):
    """SyncMoCoAugmentV5.__call__ on CUDA tensors: shapes, value ranges, flow images are k/255 levels."""
    from mscl_b200.common.ssl_aug import SyncMoCoAugmentV5
    aug = SyncMoCoAugmentV5(crop_size=112, sync_level=("batch", "batch"), t=(8, 8), flow_suffix="flow_imgs")
    torch.manual_seed(0)
    q, k = torch.rand(4, 3, 8, 112, 112).cuda(), torch.rand(4, 3, 8, 112, 112).cuda()
    aux = dict(flow_imgs_q=torch.randn(4, 2, 16, 112, 112).cuda(), flow_imgs_k=torch.randn(4, 2, 16, 112, 112).cuda())
    a, b, c = aug(q, k, aux)
    assert a.shape == q.shape and b.shape == k.shape and c["flow_imgs_q"].shape == (4, 3, 16, 112, 112)
    lv = c["flow_imgs_k"] * 255
    assert torch.all((lv - lv.round()).abs() < 1e-4) and lv.min() >= 0 and lv.max() <= 255
    assert torch.isfinite(a).all() and a.min() >= (0 - 0.485) / 0.229 - 1e-4 and a.max() <= (1 - 0.406) / 0.225 + 1e-4


def test_fetch_host_stage(fx):
    """functional.HostStage / mscl_fetch_host: small host tables reach the device through a kernel reading mapped pinned
    memory (not the copy engine) -- mixed dtypes in one launch, every slot of the ring reused several times."""
    from mscl_b200 import _cabi
    stage = fx.HostStage(slots=4)
    rng = np.random.default_rng(0)
    dev = torch.device("cuda")
    kept = []
    for it in range(13):
        a = rng.standard_normal((5 + it, 16)).astype(np.float32)
        b = rng.integers(-2 ** 40, 2 ** 40, size=7 + it)
        c = (rng.random(3 + it) < 0.5).astype(np.uint8)
        da, db, dc = stage.upload([a, b, c], dev)
        assert da.dtype == torch.float32 and db.dtype == torch.int64 and dc.dtype == torch.uint8
        assert da.data_ptr() % 16 == 0 and db.data_ptr() % 16 == 0 and dc.data_ptr() % 16 == 0
        kept.append((a, b, c, da, db, dc))
    for a, b, c, da, db, dc in kept:          # earlier uploads survive the reuse of their host slots
        np.testing.assert_array_equal(da.cpu().numpy().reshape(a.shape), a)
        np.testing.assert_array_equal(db.cpu().numpy(), b)
        np.testing.assert_array_equal(dc.cpu().numpy(), c)
    out = torch.empty(64, device=dev)
    pinned = torch.zeros(64, pin_memory=True)
    with pytest.raises(_cabi.MsclError):      # size not a multiple of 16 bytes
        _cabi.call("mscl_fetch_host", out.data_ptr(), pinned.data_ptr(), 250, torch.cuda.current_stream().cuda_stream)


# ------------------------------------------------------------------ K10 clip + SGD
def test_fused_clip_sgd_matches_torch(fx):
    """FusedClipSGD against torch.nn.utils.clip_grad_norm_(40) + torch.optim.SGD(lr=.02, momentum=.9, wd=1e-4)
    (mscl_r18_cosm_lr2e-2.py:112-119): ragged sizes, a channels_last_3d weight, a parameter without gradient,
    clipping active and inactive, four steps (first step: buf = d)."""
    from mscl_b200.optim import FusedClipSGD
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 3, 3, 7, 7), (5,), (128, 64, 3, 3, 3), (33, 17), (1,), (100003,), (16, 16, 1, 3, 3)]
    base = [torch.randn(s, generator=g) for s in shapes]
    pa = [torch.nn.Parameter(b.clone().cuda()) for b in base]
    pb = [torch.nn.Parameter(b.clone().cuda()) for b in base]
    for ps in (pa, pb):
        ps[2].data = ps[2].data.contiguous(memory_format=torch.channels_last_3d)
    ref = torch.optim.SGD(pa, lr=0.02, momentum=0.9, weight_decay=1e-4)
    opt = FusedClipSGD(pb, lr=0.02, momentum=0.9, weight_decay=1e-4, max_norm=40.0)
    for step in range(4):
        scale = (30.0, 0.01, 5.0, 0.5)[step]          # clip active on steps 0 and 2
        for i, (a, b) in enumerate(zip(pa, pb)):
            if i == 4 and step != 2:                  # a parameter the loss does not reach (grad None), most steps
                a.grad = b.grad = None
                continue
            gr = torch.randn(shapes[i], generator=g).cuda() * scale
            a.grad = gr.clone()
            b.grad = gr.clone() if i != 2 else gr.clone().contiguous(memory_format=torch.channels_last_3d)
        norm = torch.nn.utils.clip_grad_norm_(pa, 40.0)
        ref.step()
        opt.step()
        assert abs(float(opt.last_grad_norm[0]) - float(norm)) <= 1e-5 * float(norm)
        for i, (a, b) in enumerate(zip(pa, pb)):
            np.testing.assert_allclose(b.detach().cpu().numpy(), a.detach().cpu().numpy(), rtol=2e-6, atol=1e-7, err_msg=f"step {step} param {i}")
            if a.grad is not None:
                np.testing.assert_allclose(b.grad.cpu().numpy(), a.grad.cpu().numpy(), rtol=2e-6, atol=1e-7)
    sd = opt.state_dict()
    assert "momentum_buffer" in sd["state"][0]
