"""Parity of the product's recognizer-level path on the GPU -- `MSCLWithAug.objective` and a whole
`train_step` -- against the oracle and the golden fixtures produced by the unmodified reference.
Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-3     # the north star's contract: loss / gradients within 1e-3 relative (tf32 operands, fp32 accumulate)


@pytest.fixture(scope="module", autouse=True)
def _needs_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def head_level_cfg(K, t, same_kn=True, update_aug_flow=False, weight_aug_flow=(1.0, 1.0), merge_flow_epochs=True):
    """MSCLWithAug with the slim flow encoder on both branches (encoders are not run): the same model dict
    oracle/make_golden.py hands to the reference."""
    ce = dict(type="CrossEntropyLoss_torch", ignore_index=-1)

    def rec(basename):
        return dict(type="MoCoV2", backbone=dict(type="resnet_flow.r2d_18"), neck=dict(type="BaseMoCo"),
                    moco_head=dict(type="MoCoHead", basename=basename, loss_cls=ce), im_key="imgs", dim_in=128,
                    dim=128, K=K, m_base=0.994, max_iters=1000, T=0.07, mlp=True, aux_info=[],
                    aug=dict(type="IdentityAug"))

    return dict(type="MSCLWithAug", recognizer=rec(""), recognizer_flow=rec("flow"),
               moco_mx_head=dict(type="MSCLWithAugMxHead", basename="mx", loss_cls=ce, same_kn=same_kn, T=0.07),
               sup_head=dict(type="MSCLWithAugPosHeadV2", basename="", loss_pos=ce, bkb_channels=(None, None), t=t,
                             T=0.07, aux_keys=dict(im_features=dict(q_mlvl="q_mlvl"),
                                                   base_flow_features=dict(q_mlvl="q_flow_mlvl"),
                                                   aug_flow_features=dict(q_mlvl="q_aug_flow_mlvl"))),
               im_key="imgs", flow_key="flow_imgs", aux_info=[], update_aug_flow=update_aug_flow,
               weight_aug_flow=weight_aug_flow, aug=dict(type="IdentityAug"), same_kn=same_kn,
               train_cfg=dict(merge_flow_epochs=merge_flow_epochs))


def head_level_model(K, t, **switches):
    import mscl_b200
    return mscl_b200.build_model(head_level_cfg(K, t, **switches)).cuda()


def run_product_objective(inp, t, **switches):
    K = inp["queue_rgb"].shape[1]
    model = head_level_model(K, t, **switches)
    model.train()
    ptr = torch.tensor([inp["ptr"]])
    missing = model.load_state_dict({
        "recognizer.queue": inp["queue_rgb"], "recognizer.count": inp["count"], "recognizer.queue_ptr": ptr,
        "recognizer_flow.queue": inp["queue_flow"], "recognizer_flow.count": inp["count"],
        "recognizer_flow.queue_ptr": ptr}, strict=False)
    assert not missing.unexpected_keys
    leaves = {n: inp[n].cuda().requires_grad_(True) for n in ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map")}
    feats = dict(q=leaves["q"], q_f=leaves["q_f"], q_af=leaves["q_af"], k=inp["k"].cuda(), k_f=inp["k_f"].cuda(),
                 k_af=inp["k_af"].cuda(), q_mlvl=[leaves["q_map"]], q_flow_mlvl=[leaves["qf_map"]],
                 q_aug_flow_mlvl=[leaves["qaf_map"]])
    losses = model.objective(feats)
    loss, log_vars = model._parse_losses(losses)
    loss.backward()
    return model, log_vars, leaves


def _run_oracle_objective(inp, t):
    from oracle import mscl_oracle as O
    leaves = {n: inp[n].clone().requires_grad_(True) for n in ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map")}
    feats = dict(k=inp["k"], k_f=inp["k_f"], k_af=inp["k_af"], **leaves)
    rgb = O.QueueState(inp["queue_rgb"], inp["count"], inp["ptr"])
    flow = O.QueueState(inp["queue_flow"], inp["count"], inp["ptr"])
    loss, log_vars = O.parse_losses(O.mscl_objective(feats, rgb, flow, T=0.07, t=t))
    loss.backward()
    return log_vars, leaves, rgb, flow


def check_objective(inp, t, golden=None, **switches):
    """Product vs oracle (and vs the reference's golden numbers when given) for one objective call."""
    from oracle import inputs
    ref_vars, ref_leaves, rgb, flow = _run_oracle_objective(inp, t)
    model, log_vars, leaves = run_product_objective(inp, t, **switches)
    assert list(log_vars.keys()) == list(ref_vars.keys())         # the 23 keys, in the reference's order
    for k, v in log_vars.items():
        refs = [ref_vars[k]] + ([float(golden[f"logvar/{k}"])] if golden is not None else [])
        for ref in refs:
            if "acc" in k:
                assert v == pytest.approx(ref, abs=1e-6), (k, v, ref)       # integer hit counts / N
            else:
                assert abs(v - ref) <= REL * abs(ref), (k, v, ref)
    for n in ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map"):
        assert _rel(leaves[n].grad, ref_leaves[n].grad) < REL, (n, _rel(leaves[n].grad, ref_leaves[n].grad))
    if golden is not None:
        for n in ("q", "q_f", "q_af"):
            assert _rel(leaves[n].grad, torch.from_numpy(golden[f"grad/{n}"])) < REL, n
        for n in ("q_map", "qf_map", "qaf_map"):
            assert _rel(leaves[n].grad.sum(dim=(-2, -1)), torch.from_numpy(golden[f"gradsum/{n}"])) < REL, n
    # queue pointer, enqueue contents, ages: bit-exact
    for tag, rec, st in (("rgb", model.recognizer, rgb), ("flow", model.recognizer_flow, flow)):
        sd = {k: v.cpu() for k, v in rec.state_dict().items() if k in ("queue", "count", "queue_ptr")}
        assert int(sd["queue_ptr"]) == st.ptr
        np.testing.assert_array_equal(sd["queue"].numpy(), st.queue.numpy())
        np.testing.assert_array_equal(sd["count"].numpy(), st.count.numpy())
        if golden is not None:
            assert int(sd["queue_ptr"]) == int(golden[f"after/{tag}/ptr"][0])
            np.testing.assert_array_equal(sd["count"].numpy(), golden[f"after/{tag}/count"])
            if f"after/{tag}/queue" in golden.files:
                np.testing.assert_array_equal(sd["queue"].numpy(), golden[f"after/{tag}/queue"])
            else:
                assert inputs.digest(sd["queue"]) == str(golden[f"after/{tag}/queue_digest"])
    return log_vars


def test_objective_head_small_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "head_small.npz"), allow_pickle=False)
    inp = {k[3:]: (torch.from_numpy(g[k]) if g[k].ndim else int(g[k])) for k in g.files if k.startswith("in/")}
    check_objective(inp, 4, g)


def test_objective_cfg1_golden(golden_dir):
    """BASELINE config 1: N=8, C=128, K=4096, t=8 (inputs regenerated from the seed, digest-guarded)."""
    from oracle import inputs
    g = np.load(os.path.join(golden_dir, "head_cfg1.npz"), allow_pickle=False)
    kw = eval(str(g["kwargs"]))
    inp = inputs.head_inputs(**kw)
    assert inputs.digest(*[inp[k] for k in sorted(inp) if isinstance(inp[k], torch.Tensor)]) == str(g["input_digest"])
    check_objective(inp, kw["t"], g)


@pytest.mark.parametrize("N,K,t", [(32, 65536, 4), (64, 16384, 16)])
def test_objective_full_size_vs_oracle(N, K, t):
    """The r18 config's head sizes (N=32, K=65536, t=4) and BASELINE config 4's (N=64, t=16)."""
    from oracle import inputs
    inp = inputs.head_inputs(seed=5, N=N, K=K, t=t, hw_rgb=14, hw_flow=7, b_all=N)
    check_objective(inp, t)


def test_objective_three_pass_schedule_and_launch_counts(golden_dir):
    """train_cfg=dict(merge_flow_epochs=False): W_flow streamed before AND after the base-flow enqueue (three passes, two
    launches) instead of once for both queue states -- same numbers against the oracle and the reference's golden; and the
    default schedule really is ONE InfoNCE launch per step."""
    from oracle import inputs
    from mscl_b200 import _cabi
    g = np.load(os.path.join(golden_dir, "head_small.npz"), allow_pickle=False)
    inp = {k[3:]: (torch.from_numpy(g[k]) if g[k].ndim else int(g[k])) for k in g.files if k.startswith("in/")}
    check_objective(inp, 4, g, merge_flow_epochs=False)
    inp = inputs.head_inputs(seed=5, N=32, K=65536, t=4, hw_rgb=14, hw_flow=7, b_all=32)
    check_objective(inp, 4, merge_flow_epochs=False)
    names = ["mscl_infonce_fused", "mscl_infonce_fused_multi_x", "mscl_infonce_bwd_slabs", "mscl_infonce_bwd_slabs_multi", "mscl_enqueue"]
    for merge, want in ((True, [0, 1, 0, 1, 2]), (False, [1, 1, 1, 1, 2])):      # (functional.infonce_multi enters through ..._multi_x)
        _cabi.start_timing(names)
        run_product_objective(inp, 4, merge_flow_epochs=merge)
        rec = _cabi.stop_timing()
        assert [len(rec[n]) for n in names] == want, (merge, {n: len(rec[n]) for n in names})


@pytest.mark.parametrize("vname", ["cross_kn", "aug_enqueue"])
def test_objective_switches_vs_reference_golden(vname, golden_dir):
    """`same_kn=False` (rf against the RGB queue, fr against the post-enqueue flow queue) and `update_aug_flow=True,
    weight_aug_flow=(0.5, 0)` (third enqueue, halved FRA loss, no *_aug cross-modal terms) over two consecutive steps
    against the reference's numbers (tests/golden/mscl_variants.npz) and the oracle."""
    from oracle import inputs
    from test_oracle_golden import MSCL_VARIANTS, check_mscl_variant_step, run_oracle_mscl_variant
    g = np.load(os.path.join(golden_dir, "mscl_variants.npz"), allow_pickle=False)
    kw = eval(str(g["kwargs"]))
    inp = inputs.head_inputs(**kw)
    model = head_level_model(kw["K"], kw["t"], **MSCL_VARIANTS[vname]).train()
    ptr = torch.tensor([inp["ptr"]])
    model.load_state_dict({"recognizer.queue": inp["queue_rgb"], "recognizer.count": inp["count"], "recognizer.queue_ptr": ptr,
                           "recognizer_flow.queue": inp["queue_flow"], "recognizer_flow.count": inp["count"],
                           "recognizer_flow.queue_ptr": ptr}, strict=False)
    N = kw["N"]
    names = ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map")
    for step, ref_vars, ref_leaves, _, _ in run_oracle_mscl_variant(g, vname):
        x = inputs.head_inputs(**dict(kw, seed=kw["seed"] + step))
        leaves = {n: x[n].cuda().requires_grad_(True) for n in names}
        model.recognizer.note_branch(N, True)
        model.recognizer_flow.note_branch(N, True)
        model.recognizer_flow.note_branch(N, model.update_aug_flow)
        losses = model.objective(dict(q=leaves["q"], q_f=leaves["q_f"], q_af=leaves["q_af"], k=x["k"].cuda(), k_f=x["k_f"].cuda(),
                                      k_af=x["k_af"].cuda(), q_mlvl=[leaves["q_map"]], q_flow_mlvl=[leaves["qf_map"]],
                                      q_aug_flow_mlvl=[leaves["qaf_map"]]))
        loss, log_vars = model._parse_losses(losses)
        loss.backward()
        states = {}
        for br, rec in (("rgb", model.recognizer), ("flow", model.recognizer_flow)):
            sd = {k: v.cpu() for k, v in rec.state_dict().items() if k in ("queue", "count", "queue_ptr")}
            states[br] = dict(ptr=int(sd["queue_ptr"]), iters=rec.iters, count=sd["count"].numpy(), queue=sd["queue"].numpy())
        check_mscl_variant_step(g, f"{vname}/step{step}", log_vars, {n: l.grad.cpu() for n, l in leaves.items()}, states,
                                rel=REL, rel_grad=REL, rel_map=REL)
        for n in names:
            assert _rel(leaves[n].grad, ref_leaves[n].grad) < REL, (vname, step, n)


def test_objective_confident_rows():
    """Rows whose positives dominate (what training converges to): in the `rf` term the positive key and its freshly
    enqueued copy (decay 0.99999) hold almost all of the probability, and their two gradient contributions nearly
    cancel against the "- 1" of the positive.  The queue pass leaves that one column out and `finalize` adds it in
    exact fp32; with the column scored from tf32 operands the gradient of q was off by 6e-3 on such rows."""
    from oracle import inputs
    full = inputs.head_inputs(seed=11, N=32, K=4096, t=4, hw_rgb=6, hw_flow=3, b_all=8)
    inp = {k: (v[:8].contiguous() if isinstance(v, torch.Tensor) and v.shape[0] == 32 else v) for k, v in full.items()}
    check_objective(inp, 4)


def _synthetic_batch(N, T=8, S=112, seed=0):
    g = torch.Generator().manual_seed(seed)
    imgs = [torch.rand(N, 3, T, S, S, generator=g) for _ in range(2)]
    flows = [torch.rand(N, 3, 2 * T, S, S, generator=g) for _ in range(2)]
    return imgs, flows


def test_full_train_step_vs_oracle():
    """Three whole MSCLWithAug steps (R3D-18 + r2d_18 encoders, TPN neck, K=256, N=4) against the
    oracle's step on the CPU with identical weights: log vars, gradients of every trainable
    parameter, key-encoder weights after the EMA (bit-exact), queue state (pointer / ages
    bit-exact), iters / momentum bookkeeping."""
    import mscl_b200
    from mscl_b200.configs import mscl_r18_model
    from oracle.step import OracleMSCL
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        cfg = mscl_r18_model(K=256, aug="IdentityAug")
        cfg["recognizer"]["max_iters"] = cfg["recognizer_flow"]["max_iters"] = 1000
        torch.manual_seed(0)
        model = mscl_b200.build_model(cfg)
        model.train()
        orc = OracleMSCL(model)          # deep copies on the CPU, taken before the product moves
        model = model.cuda()
        N = 4
        for step in range(3):
            imgs, flows = _synthetic_batch(N, S=64, seed=10 + step)
            torch.manual_seed(100 + step)      # the shuffle permutations come from the default CPU generator
            loss_ref, vars_ref = orc.train_step(imgs[0], imgs[1], flows[0], flows[1])
            for p in orc.parameters():
                p.grad = None
            loss_ref.backward()
            torch.manual_seed(100 + step)
            model.zero_grad(set_to_none=True)
            out = model.train_step(dict(imgs=[x.cuda() for x in imgs], flow_imgs=[x.cuda() for x in flows]), None)
            out["loss"].backward()
            assert out["num_samples"] == N
            assert list(out["log_vars"].keys()) == list(vars_ref.keys())
            for k, v in out["log_vars"].items():
                if "acc" in k:
                    assert abs(v - vars_ref[k]) <= 1.0 / N + 1e-6, (step, k, v, vars_ref[k])   # near-ties may flip one row
                else:
                    assert abs(v - vars_ref[k]) <= REL * abs(vars_ref[k]), (step, k, v, vars_ref[k])
            for tag, rec, ob in (("rgb", model.recognizer, orc.rgb), ("flow", model.recognizer_flow, orc.flow)):
                assert rec.iters == ob.state.iters and rec.batch_size == ob.state.batch_size, tag
                sd = rec.state_dict()
                assert int(sd["queue_ptr"]) == ob.state.ptr
                np.testing.assert_array_equal(sd["count"].cpu().numpy(), ob.state.count.numpy())
                assert _rel(sd["queue"], ob.state.queue) < 1e-4, tag      # keys come from GPU vs CPU encoders
                # EMA of the key encoder: identical inputs (weights never stepped here) -> identical bits
                for pk, po in zip([p for m in (rec.encoder_k, rec.neck_k, rec.mlp_k) for p in m.parameters()], ob.k_params()):
                    np.testing.assert_array_equal(pk.detach().cpu().numpy(), po.detach().numpy())
            gp = [p for r in (model.recognizer, model.recognizer_flow) for m in (r.encoder_q, r.neck_q, r.mlp_q)
                  for p in m.parameters()]
            go = orc.parameters()
            assert len(gp) == len(go)
            num = sum(float((a.grad.cpu().double() - b.grad.double()).pow(2).sum()) for a, b in zip(gp, go)
                      if a.grad is not None and b.grad is not None)
            den = sum(float(b.grad.double().pow(2).sum()) for b in go if b.grad is not None)
            assert all((a.grad is None) == (b.grad is None) for a, b in zip(gp, go))
            assert (num / den) ** 0.5 < 5e-3, (step, (num / den) ** 0.5)     # through 18 conv layers, GPU vs CPU fp32
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def test_graphed_encoder_paths_match_eager():
    """mscl_b200/graphed.py: CUDA-graphed query/key encoder paths (one graph per call site, the flow recognizer is called
    twice per step) against the eager model with identical weights: log vars, gradients (incl. which parameters get
    None), key features / queue state, over three steps."""
    import mscl_b200
    from mscl_b200 import graphed
    from mscl_b200.configs import mscl_r18_model
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        cfg = mscl_r18_model(K=256, aug="IdentityAug")
        cfg["recognizer"]["max_iters"] = cfg["recognizer_flow"]["max_iters"] = 1000
        torch.manual_seed(0)
        eager = mscl_b200.build_model(cfg)
        torch.manual_seed(0)             # a second build from the same seed (a deepcopy would keep torchvision's
        fast = mscl_b200.build_model(cfg)  # patched forward bound to the original encoder)
        eager, fast = eager.train().cuda(), fast.train().cuda()
        N = 4
        imgs, flows = _synthetic_batch(N, S=64, seed=3)

        def batch(seed):
            im, fl = _synthetic_batch(N, S=64, seed=seed)
            return dict(imgs=[x.cuda() for x in im], flow_imgs=[x.cuda() for x in fl])

        def warm():     # the step graphed.enable() runs eagerly to find the parameters no loss reaches
            fast.train_step(batch(3), None)["loss"].backward()

        torch.manual_seed(50)
        state = graphed.enable(fast, imgs[0].cuda(), flows[0][:, :, :8].contiguous().cuda(), warm)
        assert len(state.unused) > 0            # some pyramid convolutions feed no loss
        torch.manual_seed(50)
        eager.train_step(batch(3), None)["loss"].backward()       # same step on the eager twin (queue, iters advance)
        for step in range(3):
            outs = []
            for m in (eager, fast):
                torch.manual_seed(200 + step)
                m.zero_grad(set_to_none=True)
                out = m.train_step(batch(20 + step), None)
                out["loss"].backward()
                outs.append(out)
            state.after_backward()
            for k, v in outs[0]["log_vars"].items():
                assert abs(v - outs[1]["log_vars"][k]) <= 2e-4 * max(1.0, abs(v)), (step, k, v, outs[1]["log_vars"][k])
            for (n, a), b in zip(eager.named_parameters(), fast.parameters()):
                assert (a.grad is None) == (b.grad is None), n
                if a.grad is not None:
                    assert _rel(b.grad, a.grad) < 2e-3, (step, n, _rel(b.grad, a.grad))
            for ra, rb in ((eager.recognizer, fast.recognizer), (eager.recognizer_flow, fast.recognizer_flow)):
                sa, sb = ra.state_dict(), rb.state_dict()
                assert int(sa["queue_ptr"]) == int(sb["queue_ptr"]) and ra.iters == rb.iters
                np.testing.assert_array_equal(sa["count"].cpu().numpy(), sb["count"].cpu().numpy())
                assert _rel(sb["queue"], sa["queue"]) < 1e-4
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def test_state_dict_roundtrip_on_device():
    """Checkpoint compatibility: reference-layout buffers out, same buffers in."""
    model = head_level_model(512, 4)
    rec = model.recognizer
    rec.negative_queue()
    keys = torch.nn.functional.normalize(torch.randn(3, 16, 128, device="cuda"), dim=2)
    for s in range(3):
        rec._dequeue_and_enqueue(keys[s])
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    assert int(sd["recognizer.queue_ptr"]) == 48 and int(sd["recognizer.count"].max()) == 3
    np.testing.assert_array_equal(sd["recognizer.queue"][:, 16:32].t().cpu().numpy(), keys[1].cpu().numpy())
    other = head_level_model(512, 4)
    other.load_state_dict(sd, strict=True)
    sd2 = other.state_dict()
    for k in sd:
        np.testing.assert_array_equal(sd2[k].cpu().numpy(), sd[k].cpu().numpy(), err_msg=k)
    np.testing.assert_allclose(other.recognizer.negative_queue().weight().cpu().numpy(),
                               (0.99999 ** sd["recognizer.count"].float().cpu() * sd["recognizer.queue"].cpu()).numpy(),
                               rtol=2e-6)
