"""Host-side logic of the sibling recognizers / heads on the CPU: which row sets are stacked against which queue state,
enqueue order, loss-dict keys and their order, 1/num_ids scaling, aux-key routing.

The CUDA kernels cannot run here, so THIS TEST substitutes the four kernel entry points the classes call
(`MoCoV2.contrast`, `MoCoV2._dequeue_and_enqueue`, `functional.hw_mean`, `functional.lmcl`) with the oracle's arithmetic
(test infrastructure only -- the product has no such path) and checks the assembled result against the golden fixtures
of the unmodified reference.  The kernels themselves are checked on the GPU (tests/test_gpu_siblings.py)."""
import os

import numpy as np
import pytest
import torch

import mscl_b200
from mscl_b200 import functional as fx
from mscl_b200.recognizers.moco import MoCoV2
from oracle import inputs, mscl_oracle as O

LOSS = dict(type="CrossEntropyLoss_torch", ignore_index=-1)


def _fake_lmcl(xq, xf, T):
    """[loss, top1, top5, 0] of normalise -> bmm -> /T -> CE vs the diagonal (what kernel `mscl_lmcl` returns)."""
    t = xq.shape[2]
    a, b = torch.nn.functional.normalize(xq, dim=1), torch.nn.functional.normalize(xf, dim=1)
    scores = torch.bmm(a.transpose(1, 2), b).flatten(0, 1) / T
    labels = torch.arange(t).repeat(xq.shape[0])
    acc = O.top_k_accuracy(scores.detach().numpy(), labels.numpy(), (1, 5))
    return torch.stack([O.cross_entropy_torch(scores, labels), torch.tensor(acc[0], dtype=torch.float32),
                        torch.tensor(acc[1], dtype=torch.float32), torch.zeros(())])


def _fake_contrast(self, terms, T=None):
    T = self.T if T is None else T
    st = self._cpu_state
    w = O.decayed_weight(st.queue, st.count)
    rows = []
    for term in terms:
        logits = O.infonce_logits(term[0], term[1].detach(), w, T)
        labels = torch.zeros(logits.shape[0], dtype=torch.long)
        acc = O.top_k_accuracy(logits.detach().numpy(), labels.numpy(), (1, 5))
        rows.append(torch.stack([O.cross_entropy_torch(logits, labels), torch.tensor(acc[0], dtype=torch.float32),
                                 torch.tensor(acc[1], dtype=torch.float32), torch.zeros(())]))
    return torch.stack(rows)


def _fake_enqueue(self, keys):
    st = self._cpu_state
    st.ptr = O.enqueue(st.queue, st.count, st.ptr, keys.detach())
    self.batch_size = keys.shape[0]


def _fake_enqueue_slots(self, n_local, device):
    return torch.arange(n_local, dtype=torch.int32) + int(self._cpu_state.ptr)


@pytest.fixture
def cpu_kernels(monkeypatch):
    monkeypatch.setattr(MoCoV2, "contrast", _fake_contrast)
    monkeypatch.setattr(MoCoV2, "_dequeue_and_enqueue", _fake_enqueue)
    monkeypatch.setattr(MoCoV2, "enqueue_slots", _fake_enqueue_slots)
    monkeypatch.setattr(fx, "hw_mean", lambda x: x.mean(dim=(-2, -1)))
    monkeypatch.setattr(fx, "lmcl", _fake_lmcl)
    monkeypatch.setattr(fx, "upsample_trilinear", lambda x, size: torch.nn.functional.interpolate(x, size=size, mode="trilinear"))


def test_sibling_heads_host_logic(cpu_kernels, golden_dir):
    g = np.load(os.path.join(golden_dir, "sibling_heads.npz"), allow_pickle=False)
    for case in inputs.sibling_head_cases():
        name, cls_name, kw, _, _, with_aug = case
        head = mscl_b200.build_head(dict(type=cls_name, basename="", loss_pos=LOSS, loss_cls=LOSS, **kw))
        head.load_state_dict({k.split("/state/")[1]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{name}/state/")},
                             strict=True)
        q_mlvl, qf_mlvl, qaf_mlvl = inputs.sibling_head_inputs(case)
        leaves = [x.requires_grad_(True) for x in q_mlvl + qf_mlvl + (qaf_mlvl or [])]
        args = dict(q_mlvl=q_mlvl, q_flow_mlvl=qf_mlvl)
        if with_aug:
            args["q_aug_flow_mlvl"] = qaf_mlvl
        losses = head.loss(**head(**args))
        sum(v for k, v in losses.items() if "loss" in k).backward()
        assert list(losses.keys()) == [str(k) for k in g[f"{name}/out_order"]]
        for k, v in losses.items():
            ref = float(g[f"{name}/out/{k}"])
            assert abs(float(v) - ref) <= 2e-6 * max(1.0, abs(ref)), (name, k, float(v), ref)
        for i, x in enumerate(leaves):
            got = x.grad.sum(dim=(-2, -1)).numpy() if x.grad is not None else np.zeros(x.shape[:3], dtype=np.float32)
            np.testing.assert_allclose(got, g[f"{name}/gradsum/{i}"], rtol=1e-4, atol=2e-6)
        for k, p in head.named_parameters():
            np.testing.assert_allclose(p.grad.numpy(), g[f"{name}/pgrad/{k}"], rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("kind", ["mscl", "modist"])
def test_two_branch_host_logic(cpu_kernels, kind, golden_dir):
    from test_gpu_siblings import _two_branch_cfg
    g = np.load(os.path.join(golden_dir, "two_branch.npz"), allow_pickle=False)
    kw = eval(str(g["kwargs"]))
    inp = inputs.head_inputs(**kw)
    model = mscl_b200.build_model(_two_branch_cfg(kind, kw["K"], kw["t"])).train()
    if kind == "mscl":
        model.sup_head.load_state_dict({k.split("/sup_state/")[1]: torch.from_numpy(g[k]) for k in g.files if "/sup_state/" in k})
    model.recognizer._cpu_state = O.QueueState(inp["queue_rgb"], inp["count"], inp["ptr"])
    model.recognizer_flow._cpu_state = O.QueueState(inp["queue_flow"], inp["count"], inp["ptr"])
    N = kw["N"]
    for step in range(2):
        x = inputs.head_inputs(**dict(kw, seed=kw["seed"] + step))
        leaves = {n: x[n].clone().requires_grad_(True) for n in ("q", "q_f", "q_map", "qf_map")}
        model.recognizer.note_branch(N, True)
        model.recognizer_flow.note_branch(N, True)
        if kind == "mscl":
            losses = model.objective(dict(q=leaves["q"], k=x["k"], q_f=leaves["q_f"], k_f=x["k_f"], aux_info={},
                                          im_features=dict(q_mlvl=[leaves["q_map"]]), flow_features=dict(q_mlvl=[leaves["qf_map"]])))
        else:
            losses = model.objective(leaves["q"], x["k"], leaves["q_f"], x["k_f"])
        loss, log_vars = model._parse_losses(losses)
        loss.backward()
        tag = f"{kind}/step{step}"
        assert list(log_vars.keys()) == [str(s) for s in g[f"{tag}/logvar_order"]]
        for key, v in log_vars.items():
            ref = float(g[f"{tag}/logvar/{key}"])
            assert abs(v - ref) <= 2e-6 * max(1.0, abs(ref)), (tag, key, v, ref)
        for n in ("q", "q_f"):
            np.testing.assert_allclose(leaves[n].grad.numpy(), g[f"{tag}/grad/{n}"], rtol=1e-5, atol=2e-6)
        for br, rec in (("rgb", model.recognizer), ("flow", model.recognizer_flow)):
            st = rec._cpu_state
            assert st.ptr == int(g[f"{tag}/after/{br}/ptr"][0]) and rec.iters == int(g[f"{tag}/after/{br}/iters"])
            np.testing.assert_array_equal(st.count.numpy(), g[f"{tag}/after/{br}/count"])
            np.testing.assert_array_equal(st.queue.numpy(), g[f"{tag}/after/{br}/queue"])


class _FakeEma:
    """fx.EmaTable stand-in: the reference's `k*m + q*(1-m)` through the oracle."""

    def __init__(self, params_k, params_q):
        self.k, self.q = list(params_k), list(params_q)

    def update(self, m):
        for pk, new in zip(self.k, O.ema_update([p.data for p in self.k], [p.data for p in self.q], m)):
            pk.data = new


def _attach_cpu_state(rec):
    st = rec._staged
    rec._cpu_state = O.QueueState(st["queue"], st["count"], st["ptr"])


@pytest.mark.parametrize("kind", ["moco", "mscl", "modist"])
def test_whole_train_step_host_logic(cpu_kernels, monkeypatch, kind):
    """forward_train plumbing of MoCo / MSCL / MoDist (EMA -> shuffle draw -> key path -> stacked terms -> enqueues ->
    iters) against the oracle's step, both on the CPU with identical weights."""
    from oracle.step import OracleMoCo, OracleTwoBranch
    from test_gpu_siblings import _moco_cfg, _two_branch_cfg
    monkeypatch.setattr(fx, "EmaTable", _FakeEma)
    torch.manual_seed(0)
    if kind == "moco":
        model = mscl_b200.build_model(_moco_cfg("MoCo", 64, m=0.99)).train()
        orc = OracleMoCo(model)
        recs = [(model, orc.branch)]
    else:
        model = mscl_b200.build_model(_two_branch_cfg(kind, 64, 4, mlvl_ids=(-1, -1))).train()
        orc = OracleTwoBranch(model, kind)
        recs = [(model.recognizer, orc.rgb), (model.recognizer_flow, orc.flow)]
    for rec, _ in recs:
        _attach_cpu_state(rec)
    N = 4
    for step in range(2):
        g = torch.Generator().manual_seed(40 + step)
        imgs = [torch.rand(N, 3, 8, 32, 32, generator=g) for _ in range(2)]
        flows = [torch.rand(N, 3, 8, 32, 32, generator=g) for _ in range(2)]
        torch.manual_seed(100 + step)
        _, vars_ref = orc.train_step(*(imgs if kind == "moco" else imgs + flows)) if kind == "moco" else \
            orc.train_step(imgs[0], imgs[1], flows[0], flows[1])
        torch.manual_seed(100 + step)
        batch = dict(imgs=imgs) if kind == "moco" else dict(imgs=imgs, flow_imgs=flows)
        out = model.train_step(batch, None)
        assert out["num_samples"] == N and list(out["log_vars"].keys()) == list(vars_ref.keys())
        for k, v in out["log_vars"].items():
            assert abs(v - vars_ref[k]) <= 5e-5 * max(1.0, abs(vars_ref[k])), (kind, step, k, v, vars_ref[k])
        for rec, ob in recs:
            assert rec._cpu_state.ptr == ob.state.ptr and rec.iters == ob.state.iters and rec.batch_size == ob.state.batch_size
            np.testing.assert_array_equal(rec._cpu_state.count.numpy(), ob.state.count.numpy())
            for pk, po in zip([p for m in (rec.encoder_k, rec.neck_k, rec.mlp_k) for p in m.parameters()], ob.k_params()):
                np.testing.assert_array_equal(pk.detach().numpy(), po.detach().numpy())


def test_parse_losses_deferred_equals_parse_losses():
    """`parse_losses_deferred` (the log-variable read postponed until `.get()`) returns exactly `_parse_losses`'
    loss tensor and values (recognizers/base.py:275-308: sum of the `loss*` means, every value averaged)."""
    from collections import OrderedDict
    from mscl_b200.recognizers import BaseMoCoRecognizer as B
    g = torch.Generator().manual_seed(0)
    losses = OrderedDict(top1_acc=torch.rand((), generator=g), loss_cls=torch.rand(4, generator=g).requires_grad_(True),
                         loss_list=[torch.rand(2, generator=g), torch.rand(3, generator=g)], top5_acc_pos=torch.rand((), generator=g))
    loss_a, vars_a = B._parse_losses(losses)
    loss_b, deferred = B.parse_losses_deferred(losses)
    assert torch.equal(loss_a, loss_b) and loss_b.requires_grad
    assert list(vars_a.keys()) == deferred.keys == ["top1_acc", "loss_cls", "loss_list", "top5_acc_pos", "loss"]
    assert vars_a == deferred.get() and deferred.get() is deferred.get()
    want = losses["loss_cls"].mean() + losses["loss_list"][0].mean() + losses["loss_list"][1].mean()
    assert abs(vars_a["loss"] - float(want)) < 1e-6


def test_r50_config_whole_step_host_logic(cpu_kernels, monkeypatch):
    """The r50 config's model (SlowOnly-R50 + TPN with one pyramid convolution, r2d_50 flow branch, LMCL head with a
    Conv1d(256,128) flow projection; `mscl_r50_cosm_lr3e-2.py`) through MSCLWithAug.train_step against the oracle's
    step -- shapes, aux-key routing, projection gradients -- on the CPU at a reduced clip size."""
    from mscl_b200.configs import mscl_r50_model
    from oracle.step import OracleMSCL
    monkeypatch.setattr(fx, "EmaTable", _FakeEma)
    cfg = mscl_r50_model(K=64, aug="IdentityAug")
    cfg["recognizer"]["max_iters"] = cfg["recognizer_flow"]["max_iters"] = 1000
    torch.manual_seed(0)
    model = mscl_b200.build_model(cfg).train()
    orc = OracleMSCL(model)
    for rec in (model.recognizer, model.recognizer_flow):
        _attach_cpu_state(rec)
    N = 2
    g = torch.Generator().manual_seed(50)
    imgs = [torch.rand(N, 3, 8, 64, 64, generator=g) for _ in range(2)]
    flows = [torch.rand(N, 3, 16, 64, 64, generator=g) for _ in range(2)]
    torch.manual_seed(100)
    loss_ref, vars_ref = orc.train_step(imgs[0], imgs[1], flows[0], flows[1])
    loss_ref.backward()
    torch.manual_seed(100)
    out = model.train_step(dict(imgs=imgs, flow_imgs=flows), None)
    out["loss"].backward()
    assert list(out["log_vars"].keys()) == list(vars_ref.keys()) and len(vars_ref) == 23
    for k, v in out["log_vars"].items():
        assert abs(v - vars_ref[k]) <= 5e-5 * max(1.0, abs(vars_ref[k])), (k, v, vars_ref[k])
    a, b = model.sup_head.trans_flow.weight.grad, orc.trans_flow.weight.grad
    assert float((a - b).norm() / b.norm()) < 1e-4
    for rec, ob in ((model.recognizer, orc.rgb), (model.recognizer_flow, orc.flow)):
        assert rec._cpu_state.ptr == ob.state.ptr and rec.iters == ob.state.iters


def _fake_enqueue_save(self, keys, save=False):
    """`_fake_enqueue` that can also hand back what it overwrote: (old keys (B,C), old counts (B,), first slot)."""
    st = self._cpu_state
    b, ptr = keys.shape[0], st.ptr
    saved = (st.queue[:, ptr:ptr + b].t().clone(), st.count[ptr:ptr + b].clone(), ptr) if save else None
    _fake_enqueue(self, keys)
    return saved


def _fake_contrast_many(calls):
    """`MoCoV2.contrast_many` with oracle arithmetic: a call that carries an epoch split (overwritten, n_pre) evaluates its
    first n_pre terms on the queue state BEFORE the recognizer's last enqueue -- rebuilt from what that enqueue saved
    (moco.py:423-440 backwards: the slots get their old keys and counts back, every other count loses one) -- and the
    others on the state as it is."""
    outs = []
    for call in calls:
        rec, terms, T = call[:3]
        split = call[3] if len(call) > 3 else None
        if split is None:
            outs.append(_fake_contrast(rec, terms, T))
            continue
        (old_keys, old_count, ptr), n_pre = split
        st = rec._cpu_state
        b = old_keys.shape[0]
        assert (ptr + b) % st.queue.shape[1] == st.ptr          # `overwritten` belongs to the LAST enqueue of this queue
        pre = O.QueueState(st.queue.clone(), st.count - 1, ptr)
        pre.queue[:, ptr:ptr + b] = old_keys.t()
        pre.count[ptr:ptr + b] = old_count
        rec._cpu_state = pre
        a = _fake_contrast(rec, terms[:n_pre], T)
        rec._cpu_state = st
        outs.append(torch.cat([a, _fake_contrast(rec, terms[n_pre:], T)]))
    return outs


@pytest.mark.parametrize("merged", [False, True])
@pytest.mark.parametrize("vname", ["cross_kn", "aug_enqueue"])
def test_mscl_with_aug_switches_host_logic(cpu_kernels, monkeypatch, vname, merged, golden_dir):
    """MSCLWithAug.objective with `same_kn=False` / `update_aug_flow=True, weight_aug_flow=(0.5, 0)`: which queue state
    each stacked term reads, the extra enqueue, the halved / dropped loss terms -- against the reference's numbers; in the
    three-pass schedule (what runs without a CUDA device) and in the one-launch schedule of the device path (`merged`: the
    base-flow enqueue first, its overwritten block handed to ONE stacked pass over W_flow, the RGB enqueue last)."""
    from test_gpu_step import head_level_cfg
    from test_oracle_golden import MSCL_VARIANTS, check_mscl_variant_step
    g = np.load(os.path.join(golden_dir, "mscl_variants.npz"), allow_pickle=False)
    kw = eval(str(g["kwargs"]))
    inp = inputs.head_inputs(**kw)
    model = mscl_b200.build_model(head_level_cfg(kw["K"], kw["t"], **MSCL_VARIANTS[vname])).train()
    launches = []
    if merged:
        monkeypatch.setattr(MoCoV2, "can_launch_together", staticmethod(lambda recs, device: True))
        monkeypatch.setattr(MoCoV2, "_dequeue_and_enqueue", _fake_enqueue_save)
        monkeypatch.setattr(MoCoV2, "contrast_many",
                            staticmethod(lambda calls: (launches.append([len(c) > 3 and c[3] is not None for c in calls]),
                                                        _fake_contrast_many(calls))[1]))
    model.recognizer._cpu_state = O.QueueState(inp["queue_rgb"], inp["count"], inp["ptr"])
    model.recognizer_flow._cpu_state = O.QueueState(inp["queue_flow"], inp["count"], inp["ptr"])
    N = kw["N"]
    names = ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map")
    for step in range(2):
        x = inputs.head_inputs(**dict(kw, seed=kw["seed"] + step))
        leaves = {n: x[n].clone().requires_grad_(True) for n in names}
        model.recognizer.note_branch(N, True)                    # what forward_train does after each encoder call
        model.recognizer_flow.note_branch(N, True)
        model.recognizer_flow.note_branch(N, model.update_aug_flow)
        losses = model.objective(dict(q=leaves["q"], q_f=leaves["q_f"], q_af=leaves["q_af"], k=x["k"], k_f=x["k_f"], k_af=x["k_af"],
                                      q_mlvl=[leaves["q_map"]], q_flow_mlvl=[leaves["qf_map"]], q_aug_flow_mlvl=[leaves["qaf_map"]]))
        loss, log_vars = model._parse_losses(losses)
        loss.backward()
        states = {br: dict(ptr=rec._cpu_state.ptr, iters=rec.iters, count=rec._cpu_state.count.numpy(), queue=rec._cpu_state.queue.numpy())
                  for br, rec in (("rgb", model.recognizer), ("flow", model.recognizer_flow))}
        check_mscl_variant_step(g, f"{vname}/step{step}", log_vars, {n: l.grad for n, l in leaves.items()}, states,
                                rel=2e-6, rel_grad=2e-5, rel_map=1e-4)
    if merged:      # per step ONE stacked launch: the W_rgb job plain, the W_flow job with the epoch split
        assert launches == [[False, True], [False, True]], launches


def test_eval_mode_step_host_logic(cpu_kernels, monkeypatch):
    """Validation runs the same train_step (eval_hooks.py:471-478): in eval() the key encoder is
    still EMA-updated, the permutation still drawn, the keys still enqueued, but `iters` (and so the momentum
    schedule) stands still (moco.py:504-505) and batch norm uses its running statistics."""
    from oracle.step import OracleMoCo
    from test_gpu_siblings import _moco_cfg
    monkeypatch.setattr(fx, "EmaTable", _FakeEma)
    torch.manual_seed(0)
    model = mscl_b200.build_model(_moco_cfg("MoCoV2", 64, m_base=0.99, max_iters=100)).train()
    orc = OracleMoCo(model)
    _attach_cpu_state(model)
    g = torch.Generator().manual_seed(70)
    batches = [[torch.rand(4, 3, 8, 32, 32, generator=g) for _ in range(2)] for _ in range(3)]
    for step, mode in enumerate((True, False, True)):          # train, validate, train
        model.train(mode)
        orc.branch.train(mode)
        orc.training = mode
        torch.manual_seed(100 + step)
        _, vars_ref = orc.train_step(*batches[step])
        torch.manual_seed(100 + step)
        out = model.train_step(dict(imgs=batches[step]), None)
        for k, v in out["log_vars"].items():
            assert abs(v - vars_ref[k]) <= 5e-5 * max(1.0, abs(vars_ref[k])), (step, k, v, vars_ref[k])
        assert model.iters == orc.branch.state.iters == (4, 4, 8)[step]
        assert model._cpu_state.ptr == orc.branch.state.ptr == 4 * (step + 1)
        assert abs(model.m - (1 - 0.5 * 0.01 * (np.cos(np.pi * (0, 0.04, 0.04)[step]) + 1))) < 1e-12
