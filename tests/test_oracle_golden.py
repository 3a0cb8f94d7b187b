"""The oracle (oracle/mscl_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py through oracle/ref_shim.py) and against the reference's own
known-answer test for top_k_accuracy.  CPU only."""
import os

import numpy as np
import torch

from oracle import inputs, mscl_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _run_oracle_head(inp, t):
    leaves = {n: inp[n].clone().requires_grad_(True) for n in ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map")}
    feats = dict(k=inp["k"], k_f=inp["k_f"], k_af=inp["k_af"], **leaves)
    rgb = O.QueueState(inp["queue_rgb"], inp["count"], inp["ptr"])
    flow = O.QueueState(inp["queue_flow"], inp["count"], inp["ptr"])
    losses = O.mscl_objective(feats, rgb, flow, T=0.07, t=t)
    loss, log_vars = O.parse_losses(losses)
    loss.backward()
    return log_vars, leaves, rgb, flow


def _check_head(g, inp, t, full_queue):
    log_vars, leaves, rgb, flow = _run_oracle_head(inp, t)
    assert list(log_vars.keys()) == [str(k) for k in g["logvar_order"]]
    for k, v in log_vars.items():
        ref = float(g[f"logvar/{k}"])
        assert abs(v - ref) <= 1e-6 * max(1.0, abs(ref)), (k, v, ref)
    for n in ("q", "q_f", "q_af"):
        np.testing.assert_allclose(leaves[n].grad.numpy(), g[f"grad/{n}"], rtol=1e-5, atol=2e-6)  # fp32 SGEMM/LSE summation order differs between hosts (thread count, ISA)
    for n in ("q_map", "qf_map", "qaf_map"):
        np.testing.assert_allclose(leaves[n].grad.sum(dim=(-2, -1)).numpy(), g[f"gradsum/{n}"], rtol=1e-4, atol=2e-6)
    for tag, st in (("rgb", rgb), ("flow", flow)):
        assert st.ptr == int(g[f"after/{tag}/ptr"][0])
        np.testing.assert_array_equal(st.count.numpy(), g[f"after/{tag}/count"])
        assert st.iters == int(g[f"after/{tag}/iters"])
        assert st.batch_size == int(g[f"after/{tag}/batch_size"])
        if full_queue:
            np.testing.assert_array_equal(st.queue.numpy(), g[f"after/{tag}/queue"])
        else:
            assert inputs.digest(st.queue) == str(g[f"after/{tag}/queue_digest"])


def test_head_small_matches_reference(golden_dir):
    g = _load(golden_dir, "head_small.npz")
    inp = {k[3:]: (torch.from_numpy(g[k]) if g[k].ndim else int(g[k])) for k in g.files if k.startswith("in/")}
    _check_head(g, inp, 4, True)


def test_head_cfg1_matches_reference(golden_dir):
    """BASELINE config 1: N=8, C=128, K=4096, t=8 -- inputs regenerated from the seed."""
    g = _load(golden_dir, "head_cfg1.npz")
    kw = eval(str(g["kwargs"]))
    inp = inputs.head_inputs(**kw)
    assert inputs.digest(*[inp[k] for k in sorted(inp) if isinstance(inp[k], torch.Tensor)]) == str(g["input_digest"]), \
        "torch CPU RNG stream differs from the build that generated the fixture"
    _check_head(g, inp, kw["t"], False)


def test_flow_iters_advance_twice(golden_dir):
    """the flow recognizer sees two forward_train calls per step."""
    g = _load(golden_dir, "head_small.npz")
    assert int(g["after/flow/iters"]) == 2 * int(g["after/rgb/iters"]) == 8


def test_enqueue_sequence(golden_dir):
    g = _load(golden_dir, "enqueue_seq.npz")
    queue = torch.from_numpy(g["queue0"]).clone()
    count = torch.zeros(queue.shape[1], dtype=torch.long)
    ptr = 0
    for s in range(g["keys"].shape[0]):
        ptr = O.enqueue(queue, count, ptr, torch.from_numpy(g["keys"][s]))
        assert ptr == int(g["ptrs"][s])
    np.testing.assert_array_equal(queue.numpy(), g["queue"])
    np.testing.assert_array_equal(count.numpy(), g["count"])
    np.testing.assert_array_equal(O.decayed_weight(queue, count).numpy(), g["weight"])


def test_ema(golden_dir):
    g = _load(golden_dir, "ema.npz")
    names = [str(n) for n in g["names"]]
    ks = [torch.from_numpy(g[f"k0/{n}"]) for n in names]
    qs = [torch.from_numpy(g[f"q/{n}"]) for n in names]
    for step, it in enumerate(g["iters"]):
        m = O.momentum(int(it), 1000, 0.994)
        assert m == float(g["m"][step])
        ks = O.ema_update(ks, qs, m)
        for n, k in zip(names, ks):
            np.testing.assert_array_equal(k.numpy(), g[f"k{step + 1}/{n}"])


def test_fra(golden_dir):
    g = _load(golden_dir, "fra.npz")
    for i in range(2):
        flows = [f for f in g[f"in{i}"]]
        out = np.stack(O.fra(flows, int(g[f"cid{i}"])))
        assert out.dtype == np.float32
        # the fixture ran under NumPy 2 (float64 rotation); the oracle restates the float32 path
        np.testing.assert_allclose(out, g[f"out{i}"], rtol=2e-6, atol=2e-7)
        T = len(flows)
        np.testing.assert_array_equal(out[:T], g[f"out{i}"][:T].astype(np.float32))   # base frames: exact


def test_shuffle(golden_dir):
    g = _load(golden_dir, "shuffle.npz")
    for seed, b in ((0, 32), (0, 128), (1234, 256)):
        torch.manual_seed(seed)
        np.testing.assert_array_equal(torch.randperm(b).numpy(), g[f"perm_seed{seed}_b{b}"])
        np.testing.assert_array_equal(torch.randperm(b).numpy(), g[f"perm2_seed{seed}_b{b}"])
    torch.manual_seed(5)
    x = torch.from_numpy(g["x"])
    idx = torch.randperm(x.shape[0])
    xs, unshuf = O.batch_shuffle(x, idx, 0, 1)
    np.testing.assert_array_equal(xs.numpy(), g["x_shuffled"])
    np.testing.assert_array_equal(unshuf.numpy(), g["idx_unshuffle"])
    np.testing.assert_array_equal(O.batch_unshuffle(xs, unshuf, 0, 1).numpy(), g["x_restored"])


def test_topk_known_answers():
    """Vectors of the reference's tests/test_metrics/test_accuracy.py:118-163."""
    scores = [np.array([-0.2203, -0.7538, 1.8789, 0.4451, -0.2526]),
              np.array([-0.0413, 0.6366, 1.1155, 0.3484, 0.0395]),
              np.array([0.0365, 0.5158, 1.1067, -0.9276, -0.2124]),
              np.array([0.6232, 0.9912, -0.8562, 0.0148, 1.6413])]
    assert O.top_k_accuracy(scores, [3, 1, 1, 1], (1,)) == [0]
    assert O.top_k_accuracy(scores, [2, 0, 4, 3], (1,)) == [0.25]
    assert O.top_k_accuracy(scores, [2, 2, 3, 1], (1,)) == [0.5]
    assert O.top_k_accuracy(scores, [2, 2, 2, 3], (1,)) == [0.75]
    assert O.top_k_accuracy(scores, [2, 2, 2, 4], (1,)) == [1.0]
    assert O.top_k_accuracy(scores, [3, 1, 1, 1], (1, 2)) == [0, 1.0]
    assert O.top_k_accuracy(scores, [3, 1, 2, 3], (1, 2)) == [0.25, 0.75]
    assert O.top_k_accuracy(scores, [1, 0, 3, 2], (1, 3, 5)) == [0, 0, 1.0]
    assert O.top_k_accuracy(scores, [1, 3, 4, 0], (1, 3, 5)) == [0, 0.5, 1.0]
    assert O.top_k_accuracy(scores, [2, 3, 0, 2], (1, 3, 5)) == [0.25, 0.75, 1.0]


def test_flow_visualizer_matches_reference(golden_dir):
    """K8 oracle vs the reference's own FlowVisualizer (common/ssl_aug.py:87-136) and colour wheel
    (tools/RAFT/core/utils/flow_viz.py:20-67), run from the reference tree by oracle/make_golden.py."""
    from oracle import mscl_oracle as O
    g = _load(golden_dir, "flowvis.npz")
    np.testing.assert_array_equal(O.colorwheel(), g["wheel"])
    out = O.flow_visualize(torch.from_numpy(g["flows"]))
    np.testing.assert_array_equal(out.numpy(), g["out"])


# ------------------------------------------------------------------ sibling heads / recognizers
def _projections(g, name, kw):
    """trans_rgb / trans_flow of a frame-level head rebuilt from the reference's stored state_dict."""
    import torch.nn as nn
    b_rgb, b_flow = kw["bkb_channels"]
    cls = str(name)
    trans_rgb = trans_flow = None
    keys = [k for k in g.files if k.startswith(f"{cls}/state/")]
    if any("trans_rgb.0.weight" in k for k in keys):
        trans_rgb = nn.Sequential(nn.Conv1d(b_rgb, 128, 1), nn.ReLU(), nn.Conv1d(128, 128, 1))
    elif any("trans_rgb.weight" in k for k in keys):
        trans_rgb = nn.Conv1d(b_rgb, 128, 1)
    if any("trans_flow.weight" in k for k in keys):
        trans_flow = nn.Conv1d(b_flow, 128, 1)
    for mod, pre in ((trans_rgb, "trans_rgb."), (trans_flow, "trans_flow.")):
        if mod is not None:
            mod.load_state_dict({k.split("/state/")[1][len(pre):]: torch.from_numpy(g[k]) for k in keys
                                 if k.split("/state/")[1].startswith(pre)}, strict=True)
    return trans_rgb, trans_flow


def run_oracle_sibling_head(g, case):
    """Oracle losses + leaves (with grads) for one sibling-head case; shared with the GPU parity test."""
    name, cls_name, kw, _, _, with_aug = case
    q_mlvl, qf_mlvl, qaf_mlvl = inputs.sibling_head_inputs(case)
    leaves = [x.requires_grad_(True) for x in q_mlvl + qf_mlvl + (qaf_mlvl or [])]
    trans_rgb, trans_flow = _projections(g, name, kw)
    if cls_name == "MlvlMSCLWithAugPosHead":
        losses = O.mlvl_lmcl(q_mlvl, qf_mlvl, qaf_mlvl, kw["mlvl_ids"], kw["mlvl_flow_ids"], kw["T"], kw["t"], trans_rgb, trans_flow)
    elif cls_name == "MoDistv2PosHead":
        losses = O.frame_contrast(q_mlvl[kw["mlvl_ids"][0]], qf_mlvl[kw["mlvl_ids"][1]], kw["T"], kw["t"], trans_rgb, trans_flow)
    else:
        losses = O.lmcl(q_mlvl[kw["mlvl_ids"][0]], qf_mlvl[kw["mlvl_ids"][1]], qaf_mlvl[kw["mlvl_ids"][1]], kw["T"], kw["t"],
                        trans_rgb, trans_flow)
    sum(v for k, v in losses.items() if "loss" in k).backward()
    params = {}
    for mod, pre in ((trans_rgb, "trans_rgb."), (trans_flow, "trans_flow.")):
        if mod is not None:
            params.update({pre + k: p for k, p in mod.named_parameters()})
    return losses, leaves, params


def test_sibling_heads_oracle_matches_reference(golden_dir):
    """frame_contrast / lmcl / mlvl_lmcl vs the reference's MSCLWithAugPosHead, MoDistv2PosHead and
    MlvlMSCLWithAugPosHead (heads/moco_head_v2.py:128-441)."""
    g = _load(golden_dir, "sibling_heads.npz")
    for case in inputs.sibling_head_cases():
        name = case[0]
        losses, leaves, params = run_oracle_sibling_head(g, case)
        assert list(losses.keys()) == [str(k) for k in g[f"{name}/out_order"]]
        for k, v in losses.items():
            ref = float(g[f"{name}/out/{k}"])
            assert abs(float(v) - ref) <= 2e-6 * max(1.0, abs(ref)), (name, k, float(v), ref)
        for i, x in enumerate(leaves):
            got = x.grad.sum(dim=(-2, -1)).numpy() if x.grad is not None else np.zeros(x.shape[:3], dtype=np.float32)
            np.testing.assert_allclose(got, g[f"{name}/gradsum/{i}"], rtol=1e-4, atol=2e-6, err_msg=f"{name} leaf {i}")
        for k, p in params.items():
            np.testing.assert_allclose(p.grad.numpy(), g[f"{name}/pgrad/{k}"], rtol=1e-4, atol=2e-6, err_msg=f"{name} {k}")


def run_oracle_two_branch(g, kind):
    """Two consecutive MSCL / MoDist steps through the oracle; yields per step (log_vars, leaves, rgb, flow)."""
    import torch.nn as nn
    kw = eval(str(g["kwargs"]))
    inp = inputs.head_inputs(**kw)
    rgb = O.QueueState(inp["queue_rgb"], inp["count"], inp["ptr"])
    flow = O.QueueState(inp["queue_flow"], inp["count"], inp["ptr"])
    trans_flow = None
    if kind == "mscl":
        trans_flow = nn.Conv1d(128, 128, 1)
        trans_flow.load_state_dict({"weight": torch.from_numpy(g["mscl/sup_state/trans_flow.weight"]),
                                    "bias": torch.from_numpy(g["mscl/sup_state/trans_flow.bias"])})
    for step in range(2):
        x = inputs.head_inputs(**dict(kw, seed=kw["seed"] + step))
        leaves = {n: x[n].clone().requires_grad_(True) for n in ("q", "q_f", "q_map", "qf_map")}
        feats = dict(k=x["k"], k_f=x["k_f"], **leaves)
        sup = None
        if kind == "mscl":
            sup = lambda: O.frame_contrast(leaves["q_map"], leaves["qf_map"], 0.07, kw["t"], None, trans_flow)
        losses = O.two_branch_objective(feats, rgb, flow, T=0.07, same_kn=True, kind=kind, sup=sup)
        loss, log_vars = O.parse_losses(losses)
        loss.backward()
        yield step, log_vars, leaves, rgb, flow


def test_two_branch_oracle_matches_reference(golden_dir):
    """two_branch_objective vs the reference's MSCL (recognizers/mscl.py:9-134) and MoDist
    (recognizers/modist.py:9-132) over two consecutive steps."""
    g = _load(golden_dir, "two_branch.npz")
    for kind in ("mscl", "modist"):
        for step, log_vars, leaves, rgb, flow in run_oracle_two_branch(g, kind):
            tag = f"{kind}/step{step}"
            assert list(log_vars.keys()) == [str(k) for k in g[f"{tag}/logvar_order"]]
            for k, v in log_vars.items():
                ref = float(g[f"{tag}/logvar/{k}"])
                assert abs(v - ref) <= 2e-6 * max(1.0, abs(ref)), (tag, k, v, ref)
            for n in ("q", "q_f"):
                np.testing.assert_allclose(leaves[n].grad.numpy(), g[f"{tag}/grad/{n}"], rtol=1e-5, atol=2e-6)
            if kind == "mscl":
                for n in ("q_map", "qf_map"):
                    np.testing.assert_allclose(leaves[n].grad.sum(dim=(-2, -1)).numpy(), g[f"{tag}/gradsum/{n}"], rtol=1e-4, atol=2e-6)
            for br, st in (("rgb", rgb), ("flow", flow)):
                assert st.ptr == int(g[f"{tag}/after/{br}/ptr"][0]) and st.iters == int(g[f"{tag}/after/{br}/iters"])
                np.testing.assert_array_equal(st.count.numpy(), g[f"{tag}/after/{br}/count"])
                np.testing.assert_array_equal(st.queue.numpy(), g[f"{tag}/after/{br}/queue"])


def test_retrieval_oracle_matches_reference(golden_dir):
    """K11 oracle vs the reference's own script lines (tools/test_retrival.py:286-304, executed by oracle/make_golden.py),
    and the rank formulation the kernel uses: acc@k == mean(#{train items above the best same-label item} < k)."""
    import torch.nn.functional as F
    g = _load(golden_dir, "retrieval.npz")
    for name in ("small", "wide"):
        kw = eval(str(g[f"{name}/kwargs"]))
        train, test, train_label, test_label = O.retrieval_inputs(**kw)
        assert inputs.digest(train, test, train_label, test_label) == str(g[f"{name}/digest"])
        acc = O.retrieval_nn_accuracy(train, test, train_label, test_label)
        np.testing.assert_allclose(acc, g[f"{name}/acc"], rtol=0, atol=1e-7)
        a = F.normalize(test - test.mean(0, keepdim=True), dim=1)
        b = F.normalize(train - train.mean(0, keepdim=True), dim=1)
        sim = a @ b.t()
        same = train_label.view(1, -1) == test_label.view(-1, 1)
        best = torch.where(same, sim, torch.full_like(sim, float("-inf"))).amax(dim=1, keepdim=True)
        rank = (sim > best).sum(dim=1)
        rank[~same.any(dim=1)] = sim.shape[1]
        by_rank = [(rank < k).float().mean().item() for k in (1, 5, 10, 20, 50)]
        np.testing.assert_allclose(by_rank, g[f"{name}/acc"], rtol=0, atol=1e-7)


MSCL_VARIANTS = {"cross_kn": dict(same_kn=False), "aug_enqueue": dict(update_aug_flow=True, weight_aug_flow=(0.5, 0.0))}


def run_oracle_mscl_variant(g, vname):
    """Two consecutive MSCLWithAug steps with non-default switches through the oracle (shared with the GPU test)."""
    kw = eval(str(g["kwargs"]))
    inp = inputs.head_inputs(**kw)
    rgb = O.QueueState(inp["queue_rgb"], inp["count"], inp["ptr"])
    flow = O.QueueState(inp["queue_flow"], inp["count"], inp["ptr"])
    names = ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map")
    for step in range(2):
        x = inputs.head_inputs(**dict(kw, seed=kw["seed"] + step))
        leaves = {n: x[n].clone().requires_grad_(True) for n in names}
        feats = dict(k=x["k"], k_f=x["k_f"], k_af=x["k_af"], **leaves)
        loss, log_vars = O.parse_losses(O.mscl_objective(feats, rgb, flow, T=0.07, t=kw["t"], **MSCL_VARIANTS[vname]))
        loss.backward()
        yield step, log_vars, leaves, rgb, flow


def check_mscl_variant_step(g, tag, log_vars, grads, states, rel, rel_grad, rel_map, exact_acc=True):
    """Compare one step's log vars / gradients / queue states with the golden entry `tag`."""
    assert list(log_vars.keys()) == [str(k) for k in g[f"{tag}/logvar_order"]]
    for k, v in log_vars.items():
        ref = float(g[f"{tag}/logvar/{k}"])
        tol = 1e-6 if ("acc" in k and exact_acc) else rel * max(1.0, abs(ref))
        assert abs(v - ref) <= tol, (tag, k, v, ref)
    for n in ("q", "q_f", "q_af"):
        a, b = grads[n].double(), torch.from_numpy(g[f"{tag}/grad/{n}"]).double()
        assert float((a - b).norm() / b.norm()) < rel_grad, (tag, n)
    for n in ("q_map", "qf_map", "qaf_map"):
        a, b = grads[n].double().sum(dim=(-2, -1)), torch.from_numpy(g[f"{tag}/gradsum/{n}"]).double()
        assert float((a - b).norm() / b.norm()) < rel_map, (tag, n)
    for br, st in states.items():
        assert st["ptr"] == int(g[f"{tag}/after/{br}/ptr"][0]) and st["iters"] == int(g[f"{tag}/after/{br}/iters"]), (tag, br)
        np.testing.assert_array_equal(st["count"], g[f"{tag}/after/{br}/count"])
        np.testing.assert_array_equal(st["queue"], g[f"{tag}/after/{br}/queue"])


def test_mscl_variants_oracle_matches_reference(golden_dir):
    """mscl_objective(same_kn=False) and (update_aug_flow=True, weight_aug_flow=(0.5, 0)) vs the reference's MSCLWithAug
    built with those switches (recognizers/mscl.py:225-277, heads/moco_head_v2.py:42-47)."""
    g = _load(golden_dir, "mscl_variants.npz")
    for vname in MSCL_VARIANTS:
        for step, log_vars, leaves, rgb, flow in run_oracle_mscl_variant(g, vname):
            states = {br: dict(ptr=st.ptr, iters=st.iters, count=st.count.numpy(), queue=st.queue.numpy())
                      for br, st in (("rgb", rgb), ("flow", flow))}
            check_mscl_variant_step(g, f"{vname}/step{step}", log_vars, {n: l.grad for n, l in leaves.items()}, states,
                                    rel=2e-6, rel_grad=2e-5, rel_map=1e-4)
