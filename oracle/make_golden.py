"""Generate tests/golden/*.npz by running the UNMODIFIED reference (through ref_shim).

    python -m oracle.make_golden          # needs /root/reference; run in the build container

The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it, the CUDA
path.  Large inputs are not stored: they are regenerated from seeds by
oracle/inputs.py and guarded by a sha1 digest stored next to the outputs.
"""
import os
import sys

import numpy as np
import torch

from . import inputs, mscl_oracle as O, ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
LOSS = dict(type="CrossEntropyLoss_torch", ignore_index=-1)


def _flow_recognizer_cfg(K, basename):
    return dict(type="MoCoV2", backbone=dict(type="resnet_flow.r2d_18"), neck=dict(type="BaseMoCo"),
                moco_head=dict(type="MoCoHead", basename=basename, loss_cls=LOSS),
                im_key="imgs", dim_in=128, dim=128, K=K, m_base=0.994, max_iters=1000, T=0.07,
                mlp=True, aux_info=[], aug=dict(type="IdentityAug"))


def build_head_level_model(ref, K, t, same_kn=True, update_aug_flow=False, weight_aug_flow=(1.0, 1.0)):
    """A real MSCLWithAug whose encoders are never run: extract_feat is patched per call."""
    cfg = dict(type="MSCLWithAug", recognizer=_flow_recognizer_cfg(K, ""),
               recognizer_flow=_flow_recognizer_cfg(K, "flow"),
               moco_mx_head=dict(type="MSCLWithAugMxHead", basename="mx", loss_cls=LOSS, same_kn=same_kn, T=0.07),
               sup_head=dict(type="MSCLWithAugPosHeadV2", basename="", loss_pos=LOSS,
                             bkb_channels=(None, None), t=t, T=0.07,
                             aux_keys=dict(im_features=dict(q_mlvl="q_mlvl"),
                                           base_flow_features=dict(q_mlvl="q_flow_mlvl"),
                                           aug_flow_features=dict(q_mlvl="q_aug_flow_mlvl"))),
               im_key="imgs", flow_key="flow_imgs", aux_info=[], update_aug_flow=update_aug_flow,
               weight_aug_flow=weight_aug_flow, aug=dict(type="SyncMoCoAugmentV5"), same_kn=same_kn)
    return ref.builder.build_model(cfg)


def run_reference_head_level(ref, inp, t, training=True):
    """Drive the reference's MSCLWithAug.train_step with given encoder outputs."""
    K = inp["queue_rgb"].shape[1]
    m = build_head_level_model(ref, K, t)
    m.train(training)
    leaves = {}
    for name in ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map"):
        leaves[name] = inp[name].clone().requires_grad_(True)
    for rec, qn in ((m.recognizer, "queue_rgb"), (m.recognizer_flow, "queue_flow")):
        rec.queue.copy_(inp[qn])
        rec.count.copy_(inp["count"])
        rec.queue_ptr[0] = inp["ptr"]
    calls_rgb = [(leaves["q"], [leaves["q_map"]], inp["k"], [], {})]
    calls_flow = [(leaves["q_f"], [leaves["qf_map"]], inp["k_f"], [], {}),
                  (leaves["q_af"], [leaves["qaf_map"]], inp["k_af"], [], {})]
    m.recognizer.extract_feat = lambda a, b: calls_rgb.pop(0)
    m.recognizer_flow.extract_feat = lambda a, b: calls_flow.pop(0)
    N = inp["q"].shape[0]
    dummy = torch.zeros(N, 3, 2, 4, 4)
    data = dict(imgs=[dummy, dummy], flow_imgs=[torch.zeros(N, 3, 4, 4, 4), torch.zeros(N, 3, 4, 4, 4)])
    out = m.train_step(data, None)
    out["loss"].backward()
    res = {f"logvar/{k}": np.float64(v) for k, v in out["log_vars"].items()}
    res["logvar_order"] = np.array(list(out["log_vars"].keys()))
    for name in ("q", "q_f", "q_af"):
        res[f"grad/{name}"] = leaves[name].grad.numpy()
    for name in ("q_map", "qf_map", "qaf_map"):
        g = leaves[name].grad
        res[f"gradsum/{name}"] = g.sum(dim=(-2, -1)).numpy()      # (N,C,t): grad is uniform over HW
        res[f"gradspread/{name}"] = np.float64((g.amax(dim=(-2, -1)) - g.amin(dim=(-2, -1))).abs().max())
    for tag, rec in (("rgb", m.recognizer), ("flow", m.recognizer_flow)):
        res[f"after/{tag}/queue"] = rec.queue.numpy().copy()
        res[f"after/{tag}/count"] = rec.count.numpy().copy()
        res[f"after/{tag}/ptr"] = rec.queue_ptr.numpy().copy()
        res[f"after/{tag}/iters"] = np.int64(rec.iters)
        res[f"after/{tag}/batch_size"] = np.int64(rec.batch_size)
    return res


def golden_head(ref, name, store_inputs, **kw):
    t = kw["t"]
    inp = inputs.head_inputs(**kw)
    res = run_reference_head_level(ref, inp, t)
    res["input_digest"] = np.array(inputs.digest(*[inp[k] for k in sorted(inp) if isinstance(inp[k], torch.Tensor)]))
    res["kwargs"] = np.array(repr(kw))
    if store_inputs:
        for k, v in inp.items():
            res[f"in/{k}"] = v.numpy() if isinstance(v, torch.Tensor) else np.int64(v)
    else:  # queues after enqueue are big: keep only the written block and a digest
        for tag in ("rgb", "flow"):
            q = res.pop(f"after/{tag}/queue")
            p, b = inp["ptr"], kw["N"]
            res[f"after/{tag}/queue_block"] = q[:, p:p + b].copy()
            res[f"after/{tag}/queue_digest"] = np.array(inputs.digest(q))
    np.savez_compressed(os.path.join(OUT, name), **res)
    print(name, {k: round(float(v), 6) for k, v in res.items() if k.startswith("logvar/")})


def golden_enqueue(ref):
    """Six enqueues of 16 keys into K=64 (wraps) through the real _dequeue_and_enqueue."""
    m = ref.builder.build_recognizer(_flow_recognizer_cfg(64, ""))
    g = torch.Generator().manual_seed(7)
    q0 = m.queue.clone()
    keys = torch.randn(6, 16, 128, generator=g)
    ptrs = []
    for s in range(6):
        m._dequeue_and_enqueue(keys[s])
        ptrs.append(int(m.queue_ptr))
    np.savez_compressed(os.path.join(OUT, "enqueue_seq.npz"), queue0=q0.numpy(), keys=keys.numpy(),
                        queue=m.queue.numpy(), count=m.count.numpy(), ptrs=np.array(ptrs),
                        weight=(0.99999 ** (1.0 * m.count) * m.queue).numpy())


def golden_ema(ref):
    """Three real _momentum_update_key_encoder calls on the flow recognizer's MLP + stem."""
    torch.manual_seed(3)
    m = ref.builder.build_recognizer(_flow_recognizer_cfg(64, ""))
    g = torch.Generator().manual_seed(11)
    for p in m.parameters():
        p.data.add_(torch.randn(p.shape, generator=g) * 0.05)
    names = ["mlp_k.0.weight", "mlp_k.0.bias", "mlp_k.2.weight", "encoder_k.stem.0.weight",
             "encoder_k.layer4.1.conv2.1.bias"]
    sd = dict(m.named_parameters())
    res = {}
    for n in names:
        res[f"k0/{n}"] = sd[n].detach().numpy().copy()
        res[f"q/{n}"] = sd[n.replace("_k", "_q")].detach().numpy().copy()
    ms = []
    m.max_iters = 1000
    for step, iters in enumerate((0, 250, 1000)):
        m.iters = iters
        m._momentum_update_key_encoder()
        ms.append(m.m)
        sd = dict(m.named_parameters())
        for n in names:
            res[f"k{step + 1}/{n}"] = sd[n].detach().numpy().copy()
    res["m"] = np.array(ms, dtype=np.float64)
    res["iters"] = np.array([0, 250, 1000])
    res["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "ema.npz"), **res)
    print("ema m:", ms)


def golden_fra(ref):
    """Real NormFlowWithStidedAug on two small clips (cid from np.random.seed)."""
    res = {}
    for i, seed in enumerate((0, 5)):
        flows = inputs.flow_clip(seed=seed, T=4, H=16, W=24)
        np.random.seed(seed)
        tr = ref.NormFlowWithStidedAug(ratios=(0.2, 1.8), num_chunks=8, merge_aug=True)
        out = tr(dict(flows=[f.copy() for f in flows]))
        res[f"in{i}"] = np.stack(flows)
        res[f"out{i}"] = np.stack(out["flow_imgs"]).astype(np.float64)
        res[f"cid{i}"] = np.int64(out["ap_labels"])
    np.savez_compressed(os.path.join(OUT, "fra.npz"), **res)
    print("fra cids:", res["cid0"], res["cid1"])


def golden_shuffle(ref):
    """The permutation source: torch.randperm on the CPU default generator (moco.py:160)."""
    res = {}
    for seed, b in ((0, 32), (0, 128), (1234, 256)):
        torch.manual_seed(seed)
        a = torch.randperm(b)
        c = torch.randperm(b)   # second draw of the same step (flow branch)
        res[f"perm_seed{seed}_b{b}"] = a.numpy()
        res[f"perm2_seed{seed}_b{b}"] = c.numpy()
    # the real single-rank shuffle/unshuffle round trip
    m = ref.builder.build_recognizer(_flow_recognizer_cfg(64, ""))
    torch.manual_seed(5)
    x = torch.arange(8 * 3, dtype=torch.float32).view(8, 3)
    xs, unshuf = m._batch_shuffle_ddp(x)
    res["x"], res["x_shuffled"], res["idx_unshuffle"] = x.numpy(), xs.numpy(), unshuf.numpy()
    res["x_restored"] = m._batch_unshuffle_ddp(xs, unshuf).numpy()
    np.savez_compressed(os.path.join(OUT, "shuffle.npz"), **res)


def golden_flowvis(ref):
    """The reference's own FlowVisualizer / flow_uv_to_colors (common/ssl_aug.py:87-136) with the reference's colour
    wheel (tools/RAFT/core/utils/flow_viz.py:20-67), executed from the reference tree; kornia is only needed by the
    other definitions of that file, so just these three are loaded."""
    import math
    viz = ref_shim._load("mscl_ref_flow_viz", "tools/RAFT/core/utils/flow_viz.py")
    defs = ref_shim.load_defs("mmaction/models/common/ssl_aug.py", ["flow_uv_to_colors", "FlowVisualizer"],
                              dict(torch=torch, math=math, make_colorwheel=viz.make_colorwheel))
    g = torch.Generator().manual_seed(7)
    flows = torch.randn(2, 2, 3, 12, 16, generator=g) * torch.tensor([0.4, 1.5]).view(2, 1, 1, 1, 1)
    flows[0, :, 0, 0, :4] = torch.tensor([[0.0, 1.0, -1.0, 0.0], [0.0, 0.0, 0.0, -1.0]])   # axis-aligned / zero vectors
    out = defs.FlowVisualizer()(flows.clone())
    np.savez_compressed(os.path.join(OUT, "flowvis.npz"), flows=flows.numpy(), out=out.numpy(),
                        wheel=viz.make_colorwheel())
    print("flowvis:", tuple(out.shape), out.dtype, float(out.mean()))


class _AttrDict(dict):
    """Enough of mmcv's ConfigDict for modist.py:43-44 (`moco_head.copy(); moco_head_r.basename += '_r'`)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def copy(self):
        return _AttrDict(self)


def golden_sibling_heads(ref):
    """MSCLWithAugPosHead, MoDistv2PosHead, MlvlMSCLWithAugPosHead of the unmodified reference
    (heads/moco_head_v2.py:128-441): weights, losses, accuracies and input gradients."""
    import sys
    mod = sys.modules["mmaction.models.heads.moco_head_v2"]
    res = {}
    for case in inputs.sibling_head_cases():
        name, cls_name, kw, _, _, with_aug = case
        torch.manual_seed(11)
        head = getattr(mod, cls_name)(basename="", loss_pos=LOSS, loss_cls=LOSS, **kw)
        q_mlvl, qf_mlvl, qaf_mlvl = inputs.sibling_head_inputs(case)
        leaves = [x.requires_grad_(True) for x in q_mlvl + qf_mlvl + (qaf_mlvl or [])]
        args = dict(q_mlvl=q_mlvl, q_flow_mlvl=qf_mlvl)
        if with_aug:
            args["q_aug_flow_mlvl"] = qaf_mlvl
        out = head(**args)
        losses = head.loss(**out)
        total = sum(v for k, v in losses.items() if "loss" in k)
        total.backward()
        for k, v in head.state_dict().items():
            res[f"{name}/state/{k}"] = v.detach().numpy().copy()
        for k, v in losses.items():
            res[f"{name}/out/{k}"] = np.float64(float(v))
        res[f"{name}/out_order"] = np.array(list(losses.keys()))
        for i, x in enumerate(leaves):
            res[f"{name}/gradsum/{i}"] = (x.grad.sum(dim=(-2, -1)).numpy() if x.grad is not None
                                          else np.zeros(x.shape[:3], dtype=np.float32))
        for k, p_ in head.named_parameters():
            res[f"{name}/pgrad/{k}"] = p_.grad.numpy().copy()
        print(name, {k: round(float(v), 6) for k, v in losses.items()})
    np.savez_compressed(os.path.join(OUT, "sibling_heads.npz"), **res)


def _two_branch_model(ref, kind, K, t, mlvl_ids=(0, -1)):
    import sys
    if kind == "mscl":
        cfg = dict(type="MSCL", recognizer=_flow_recognizer_cfg(K, ""), recognizer_flow=_flow_recognizer_cfg(K, "flow"),
                   moco_mx_head=dict(type="MSCLWithAugMxHead", basename="mx", loss_cls=LOSS, same_kn=True, T=0.07),
                   sup_head=dict(type="MoDistv2PosHead", basename="", loss_pos=LOSS, bkb_channels=(None, 128), t=t, T=0.07,
                                 mlvl_ids=mlvl_ids, aux_keys=dict(im_features=dict(q_mlvl="q_mlvl"),
                                               base_flow_features=dict(q_mlvl="q_flow_mlvl"))),
                   im_key="imgs", flow_key="flow_imgs", flow_img_key="flow_imgs", aux_info=[], aug=dict(type="IdentityAug"),
                   same_kn=True)
        return ref.builder.build_model(cfg)
    if "mmaction.models.recognizers.modist" not in sys.modules:
        from . import ref_shim
        ref_shim._load("mmaction.models.recognizers.modist", "mmaction/models/recognizers/modist.py")
    cfg = dict(type="MoDist", recognizer=_flow_recognizer_cfg(K, ""), recognizer_flow=_flow_recognizer_cfg(K, "flow"),
               moco_head=_AttrDict(type="MoCoHead", basename="mx", loss_cls=LOSS), im_key="imgs", flow_key="flow_imgs",
               aux_info=[], aug=dict(type="IdentityAug"), same_kn=True)
    m = ref.builder.build_model(cfg)
    m.aug_gpu.forward_with_flow = lambda a, b, c, d, e: (a, b, c, d, e)
    return m


def golden_two_branch(ref):
    """MSCL (with a MoDistv2PosHead) and MoDist of the unmodified reference (recognizers/mscl.py:9-134,
    recognizers/modist.py:9-132), driven at head level like golden_head: encoder outputs given, two consecutive
    steps so that the second one sees non-trivial ages / pointer / iters."""
    kw = dict(seed=2, N=4, C=128, K=256, t=4, hw_rgb=6, hw_flow=3)
    res = {"kwargs": np.array(repr(kw))}
    for kind in ("mscl", "modist"):
        inp = inputs.head_inputs(**kw)
        torch.manual_seed(13)
        m = _two_branch_model(ref, kind, kw["K"], kw["t"])
        m.train()
        for rec, qn in ((m.recognizer, "queue_rgb"), (m.recognizer_flow, "queue_flow")):
            rec.queue.copy_(inp[qn])
            rec.count.copy_(inp["count"])
            rec.queue_ptr[0] = inp["ptr"]
        if kind == "mscl":
            for k, v in m.sup_head.state_dict().items():
                res[f"mscl/sup_state/{k}"] = v.numpy().copy()
        N = kw["N"]
        dummy = torch.zeros(N, 3, 2, 4, 4)
        for step in range(2):
            x = inputs.head_inputs(**dict(kw, seed=kw["seed"] + step))
            leaves = {n: x[n].clone().requires_grad_(True) for n in ("q", "q_f", "q_map", "qf_map")}
            calls_rgb = [(leaves["q"], [leaves["q_map"]], x["k"], [], {})]
            calls_flow = [(leaves["q_f"], [leaves["qf_map"]], x["k_f"], [], {})]
            m.recognizer.extract_feat = lambda a, b: calls_rgb.pop(0)
            m.recognizer_flow.extract_feat = lambda a, b: calls_flow.pop(0)
            out = m.train_step(dict(imgs=[dummy, dummy], flow_imgs=[dummy, dummy]), None)
            out["loss"].backward()
            tag = f"{kind}/step{step}"
            for k, v in out["log_vars"].items():
                res[f"{tag}/logvar/{k}"] = np.float64(v)
            res[f"{tag}/logvar_order"] = np.array(list(out["log_vars"].keys()))
            for n in ("q", "q_f"):
                res[f"{tag}/grad/{n}"] = leaves[n].grad.numpy()
            if kind == "mscl":
                for n in ("q_map", "qf_map"):
                    res[f"{tag}/gradsum/{n}"] = leaves[n].grad.sum(dim=(-2, -1)).numpy()
            for br, rec in (("rgb", m.recognizer), ("flow", m.recognizer_flow)):
                res[f"{tag}/after/{br}/queue"] = rec.queue.numpy().copy()
                res[f"{tag}/after/{br}/count"] = rec.count.numpy().copy()
                res[f"{tag}/after/{br}/ptr"] = rec.queue_ptr.numpy().copy()
                res[f"{tag}/after/{br}/iters"] = np.int64(rec.iters)
            print(tag, {k: round(float(v), 5) for k, v in out["log_vars"].items()})
    np.savez_compressed(os.path.join(OUT, "two_branch.npz"), **res)


def golden_retrieval(ref):
    """Lines 286-304 of the reference's tools/test_retrival.py (the script body after feature extraction), executed
    from the reference tree on seeded features: the five kNN accuracies."""
    import textwrap
    import torch.nn.functional as F
    path = os.path.join(ref_shim.REF_ROOT, "tools", "test_retrival.py")
    with open(path) as f:
        lines = f.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip().startswith("ks = [1,5,10,20,50]"))
    end = next(i for i, l in enumerate(lines) if i > start and l.strip().startswith("print('%dNN acc"))
    body = textwrap.dedent("\n".join(lines[start:end + 1]))
    res = {}
    for name, kw in (("small", dict(seed=0)), ("wide", dict(seed=1, n_train=1500, n_test=400, dim=512, n_classes=101, noise=6.0))):
        train, test, train_label, test_label = O.retrieval_inputs(**kw)
        ns = dict(torch=torch, F=F, train_feature=train.clone(), test_feature=test.clone(), train_label=train_label,
                  test_label=test_label)
        exec(compile(body, path, "exec"), ns)
        res[f"{name}/acc"] = np.array(ns["NN_acc"], dtype=np.float64)
        res[f"{name}/kwargs"] = np.array(repr(kw))
        res[f"{name}/digest"] = np.array(inputs.digest(train, test, train_label, test_label))
        print("retrieval", name, ns["NN_acc"])
    np.savez_compressed(os.path.join(OUT, "retrieval.npz"), **res)


MSCL_VARIANTS = {"cross_kn": dict(same_kn=False),
                 "aug_enqueue": dict(update_aug_flow=True, weight_aug_flow=(0.5, 0.0))}


def golden_mscl_variants(ref):
    """MSCLWithAug with the switches the r18 / r50 configs leave at their defaults (recognizers/mscl.py:225-277,
    heads/moco_head_v2.py:42-47): `same_kn=False` (rf against the RGB queue, fr against the flow queue) and
    `update_aug_flow=True, weight_aug_flow=(0.5, 0)` (the FRA-flow call enqueues too, its loss is halved, no *_aug
    cross-modal terms).  Two consecutive head-level steps each."""
    kw = dict(seed=4, N=4, C=128, K=256, t=4, hw_rgb=6, hw_flow=3)
    res = {"kwargs": np.array(repr(kw))}
    names = ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map")
    for vname, opts in MSCL_VARIANTS.items():
        inp = inputs.head_inputs(**kw)
        m = build_head_level_model(ref, kw["K"], kw["t"], **opts)
        m.train()
        for rec, qn in ((m.recognizer, "queue_rgb"), (m.recognizer_flow, "queue_flow")):
            rec.queue.copy_(inp[qn])
            rec.count.copy_(inp["count"])
            rec.queue_ptr[0] = inp["ptr"]
        N = kw["N"]
        dummy = torch.zeros(N, 3, 2, 4, 4)
        for step in range(2):
            x = inputs.head_inputs(**dict(kw, seed=kw["seed"] + step))
            leaves = {n: x[n].clone().requires_grad_(True) for n in names}
            calls_rgb = [(leaves["q"], [leaves["q_map"]], x["k"], [], {})]
            calls_flow = [(leaves["q_f"], [leaves["qf_map"]], x["k_f"], [], {}),
                          (leaves["q_af"], [leaves["qaf_map"]], x["k_af"], [], {})]
            m.recognizer.extract_feat = lambda a, b: calls_rgb.pop(0)
            m.recognizer_flow.extract_feat = lambda a, b: calls_flow.pop(0)
            out = m.train_step(dict(imgs=[dummy, dummy], flow_imgs=[torch.zeros(N, 3, 4, 4, 4)] * 2), None)
            out["loss"].backward()
            tag = f"{vname}/step{step}"
            for k, v in out["log_vars"].items():
                res[f"{tag}/logvar/{k}"] = np.float64(v)
            res[f"{tag}/logvar_order"] = np.array(list(out["log_vars"].keys()))
            for n in ("q", "q_f", "q_af"):
                res[f"{tag}/grad/{n}"] = leaves[n].grad.numpy()
            for n in ("q_map", "qf_map", "qaf_map"):
                res[f"{tag}/gradsum/{n}"] = leaves[n].grad.sum(dim=(-2, -1)).numpy()
            for br, rec in (("rgb", m.recognizer), ("flow", m.recognizer_flow)):
                res[f"{tag}/after/{br}/queue"] = rec.queue.numpy().copy()
                res[f"{tag}/after/{br}/count"] = rec.count.numpy().copy()
                res[f"{tag}/after/{br}/ptr"] = rec.queue_ptr.numpy().copy()
                res[f"{tag}/after/{br}/iters"] = np.int64(rec.iters)
            print(tag, {k: round(float(v), 5) for k, v in out["log_vars"].items()})
    np.savez_compressed(os.path.join(OUT, "mscl_variants.npz"), **res)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load_reference()
    ref_shim.ensure_process_group()
    golden_head(ref, "head_small.npz", True, seed=1, N=4, C=128, K=256, t=4, hw_rgb=6, hw_flow=3)
    golden_head(ref, "head_cfg1.npz", False, seed=0, N=8, C=128, K=4096, t=8, hw_rgb=28, hw_flow=7)
    golden_enqueue(ref)
    golden_ema(ref)
    golden_fra(ref)
    golden_shuffle(ref)
    golden_flowvis(ref)
    golden_sibling_heads(ref)
    golden_two_branch(ref)
    golden_retrieval(ref)
    golden_mscl_variants(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.exit(main())
