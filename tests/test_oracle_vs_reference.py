"""Live pin: the oracle's full training step (oracle/step.py, which re-uses the product's
PyTorch encoder / neck modules on the CPU) against the UNMODIFIED reference executed through
oracle/ref_shim.py.  Also proves state_dict key compatibility of the product model with the
reference model.  Skipped where /root/reference does not exist (the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def pair():
    import mscl_b200
    from mscl_b200.configs import mscl_r18_model
    ref = ref_shim.load_reference()
    ref_shim.ensure_process_group()
    cfg = mscl_r18_model(K=256, aug="IdentityAug")
    cfg["recognizer"]["max_iters"] = cfg["recognizer_flow"]["max_iters"] = 1000
    torch.manual_seed(0)
    ref_model = ref.builder.build_model(dict(cfg, aug=dict(type="SyncMoCoAugmentV5")))
    mine = mscl_b200.build_model(cfg)
    return ref, ref_model, mine


def test_state_dict_keys_and_load(pair):
    ref, ref_model, mine = pair
    sd_ref = ref_model.state_dict()
    sd_mine = mine.state_dict()
    assert set(sd_ref.keys()) == set(sd_mine.keys())
    for k in sd_ref:
        assert tuple(sd_ref[k].shape) == tuple(sd_mine[k].shape), k
        assert sd_ref[k].dtype == sd_mine[k].dtype, k
    mine.load_state_dict(sd_ref, strict=True)
    sd2 = mine.state_dict()
    for k in sd_ref:
        np.testing.assert_array_equal(sd2[k].numpy(), sd_ref[k].numpy(), err_msg=k)


def test_parameter_counts(pair):
    """SURVEY App. B probe facts."""
    _, _, mine = pair
    count = lambda ms: sum(p.numel() for m in ms for p in m.parameters())
    r, f = mine.recognizer, mine.recognizer_flow
    assert count([r.encoder_k, r.neck_k, r.mlp_k]) == 36_707_392
    assert count([f.encoder_k, f.neck_k, f.mlp_k]) == 735_120
    assert len([p for m in (r.encoder_k, r.neck_k, r.mlp_k) for p in m.parameters()]) == 88
    assert len([p for m in (f.encoder_k, f.neck_k, f.mlp_k) for p in m.parameters()]) == 64


def test_full_step_matches_reference(pair):
    from oracle.step import OracleMSCL
    ref, ref_model, mine = pair
    mine.load_state_dict(ref_model.state_dict(), strict=True)
    ref_model.train(), mine.train()
    orc = OracleMSCL(mine)
    g = torch.Generator().manual_seed(3)
    N = 2
    data = dict(imgs=[torch.rand(N, 3, 8, 112, 112, generator=g) for _ in range(2)],
                flow_imgs=[torch.rand(N, 3, 16, 112, 112, generator=g) for _ in range(2)])
    for step in range(2):      # second step exercises non-zero iters / ages / pointer
        torch.manual_seed(100 + step)
        out = ref_model.train_step(data, None)
        ref_model.zero_grad()
        out["loss"].backward()
        torch.manual_seed(100 + step)
        for p in orc.parameters():
            p.grad = None
        loss, log_vars = orc.train_step(data["imgs"][0], data["imgs"][1], data["flow_imgs"][0], data["flow_imgs"][1])
        loss.backward()
        assert list(log_vars) == list(out["log_vars"])
        for k, v in out["log_vars"].items():
            assert abs(log_vars[k] - v) <= 2e-5 * max(1.0, abs(v)), (step, k, log_vars[k], v)
        for tag, rec, st in (("rgb", ref_model.recognizer, orc.rgb.state), ("flow", ref_model.recognizer_flow, orc.flow.state)):
            assert int(rec.queue_ptr) == st.ptr and rec.iters == st.iters and rec.batch_size == st.batch_size
            np.testing.assert_array_equal(rec.count.numpy(), st.count.numpy())
            np.testing.assert_allclose(rec.queue.numpy(), st.queue.numpy(), rtol=0, atol=2e-6)
        gr = dict(ref_model.named_parameters())
        for name, p in (("recognizer.mlp_q.2.weight", orc.rgb.mlp_q[2].weight),
                        ("recognizer_flow.encoder_q.stem.0.weight", orc.flow.encoder_q.stem[0].weight),
                        ("recognizer.neck_q.tpn.fpn.lateral_convs.0.conv.weight", orc.rgb.neck_q.tpn.fpn.lateral_convs[0].conv.weight)):
            a, b = p.grad, gr[name].grad
            assert float((a - b).norm() / b.norm()) < 2e-4, name
        # key encoders after the EMA(s)
        kr = dict(ref_model.named_parameters())
        np.testing.assert_allclose(orc.flow.mlp_k[0].weight.detach().numpy(), kr["recognizer_flow.mlp_k.0.weight"].detach().numpy(),
                                   rtol=0, atol=1e-7)


def test_fra_live(pair):
    """The oracle's float32 FRA against the reference's NumPy code run here (NumPy 2: float64 rotation)."""
    from oracle import inputs, mscl_oracle as O
    ref = pair[0]
    flows = inputs.flow_clip(seed=4, T=8, H=32, W=32)
    np.random.seed(11)
    out = ref.NormFlowWithStidedAug(ratios=(0.2, 1.8), num_chunks=8)(dict(flows=[f.copy() for f in flows]))
    got = np.stack(O.fra(flows, int(out["ap_labels"])))
    np.testing.assert_allclose(got, np.stack(out["flow_imgs"]), rtol=3e-6, atol=3e-7)


# ------------------------------------------------------------------ sibling recognizers / backbone
def test_slowonly_backbone_matches_reference():
    """mscl_b200's ResNet3dSlowOnly against the reference class with the r50 configs' arguments: state_dict keys,
    parameter count of BASELINE config 5 (31,672,128 in 159 tensors), identical init stream and forward outputs."""
    from mscl_b200.backbones import ResNet3dSlowOnly
    Ref = ref_shim.load_slowonly()
    kw = dict(depth=50, pretrained=None, pretrained2d=False, lateral=False, num_stages=4, conv1_kernel=(5, 7, 7),
              conv1_stride_t=2, pool1_stride_t=1, spatial_strides=(1, 2, 2, 2), out_indices=(0, 1, 2, 3))
    r, m = Ref(**kw), ResNet3dSlowOnly(**kw)
    torch.manual_seed(1)
    r.init_weights()
    torch.manual_seed(1)
    m.init_weights()
    sr, sm = r.state_dict(), m.state_dict()
    assert list(sr) == list(sm)
    for k in sr:
        assert torch.equal(sr[k], sm[k]), k
    assert sum(p.numel() for p in m.parameters()) == 31_672_128 and len(list(m.parameters())) == 159
    for k, v in sr.items():     # the zero-initialised last batch norms would hide the residual branches
        if ".bn." in k and k.endswith(("weight", "bias")):
            v.uniform_(0.5, 1.5)
    m.load_state_dict(sr, strict=True)
    x = torch.rand(2, 3, 8, 64, 64)
    for mode in (False, True):
        r.train(mode), m.train(mode)
        for a, b in zip(r(x), m(x)):
            assert a.shape == b.shape and torch.equal(a, b)


def _small_moco_cfg(typ, K, **kw):
    cfg = dict(type=typ, backbone=dict(type="resnet_flow.r2d_18"), neck=dict(type="BaseMoCo"),
               moco_head=dict(type="MoCoHead", basename="", loss_cls=dict(type="CrossEntropyLoss_torch", ignore_index=-1)),
               im_key="imgs", dim_in=128, dim=128, K=K, T=0.07, mlp=True, aux_info=[], aug=dict(type="IdentityAug"))
    cfg.update(kw)
    return cfg


def test_moco_v1_full_step_matches_reference():
    """`MoCo` (constant momentum, recognizers/moco.py:30-316) -- oracle step vs the unmodified reference, 2 steps."""
    import mscl_b200
    from oracle.step import OracleMoCo
    ref = ref_shim.load_reference()
    ref_shim.ensure_process_group()
    cfg = _small_moco_cfg("MoCo", 64, m=0.99)
    torch.manual_seed(0)
    ref_model = ref.builder.build_model(cfg)
    mine = mscl_b200.build_model(cfg)
    assert set(ref_model.state_dict()) == set(mine.state_dict())
    mine.load_state_dict(ref_model.state_dict(), strict=True)
    ref_model.train(), mine.train()
    orc = OracleMoCo(mine)
    g = torch.Generator().manual_seed(5)
    data = dict(imgs=[torch.rand(4, 3, 8, 32, 32, generator=g) for _ in range(2)])
    for step in range(2):
        torch.manual_seed(100 + step)
        out = ref_model.train_step(data, None)
        ref_model.zero_grad()
        out["loss"].backward()
        torch.manual_seed(100 + step)
        for p in orc.parameters():
            p.grad = None
        loss, log_vars = orc.train_step(data["imgs"][0], data["imgs"][1])
        loss.backward()
        assert list(log_vars) == list(out["log_vars"])
        for k, v in out["log_vars"].items():
            assert abs(log_vars[k] - v) <= 2e-5 * max(1.0, abs(v)), (step, k, log_vars[k], v)
        st = orc.branch.state
        assert int(ref_model.queue_ptr) == st.ptr
        np.testing.assert_array_equal(ref_model.count.numpy(), st.count.numpy())
        np.testing.assert_allclose(ref_model.queue.numpy(), st.queue.numpy(), rtol=0, atol=2e-6)
        kr = dict(ref_model.named_parameters())
        np.testing.assert_allclose(orc.branch.mlp_k[0].weight.detach().numpy(), kr["mlp_k.0.weight"].detach().numpy(), rtol=0, atol=1e-7)
        a, b = orc.branch.mlp_q[2].weight.grad, kr["mlp_q.2.weight"].grad
        assert float((a - b).norm() / b.norm()) < 2e-4


@pytest.mark.parametrize("kind", ["mscl", "modist"])
def test_two_branch_full_step_matches_reference(kind):
    """MSCL (+ MoDistv2PosHead) and MoDist: oracle step (oracle/step.py OracleTwoBranch) vs the unmodified reference."""
    import mscl_b200
    from oracle import make_golden as G
    from oracle.step import OracleTwoBranch
    ref = ref_shim.load_reference()
    ref_shim.ensure_process_group()
    K, t = 64, 4
    torch.manual_seed(0)
    ref_model = G._two_branch_model(ref, kind, K, t, mlvl_ids=(-1, -1))
    if kind == "mscl":
        cfg = dict(type="MSCL", recognizer=G._flow_recognizer_cfg(K, ""), recognizer_flow=G._flow_recognizer_cfg(K, "flow"),
                   moco_mx_head=dict(type="MSCLWithAugMxHead", basename="mx", loss_cls=G.LOSS, same_kn=True, T=0.07),
                   sup_head=dict(type="MoDistv2PosHead", basename="", loss_pos=G.LOSS, bkb_channels=(None, 128), t=t, T=0.07,
                                 mlvl_ids=(-1, -1), aux_keys=dict(im_features=dict(q_mlvl="q_mlvl"), base_flow_features=dict(q_mlvl="q_flow_mlvl"))),
                   im_key="imgs", flow_key="flow_imgs", flow_img_key="flow_imgs", aux_info=[], aug=dict(type="IdentityAug"), same_kn=True)
    else:
        cfg = dict(type="MoDist", recognizer=G._flow_recognizer_cfg(K, ""), recognizer_flow=G._flow_recognizer_cfg(K, "flow"),
                   moco_head=dict(type="MoCoHead", basename="mx", loss_cls=G.LOSS), im_key="imgs", flow_key="flow_imgs",
                   aux_info=[], aug=dict(type="IdentityAug"), same_kn=True)
    mine = mscl_b200.build_model(cfg)
    assert set(ref_model.state_dict()) == set(mine.state_dict())
    mine.load_state_dict(ref_model.state_dict(), strict=True)
    ref_model.train(), mine.train()
    orc = OracleTwoBranch(mine, kind)
    g = torch.Generator().manual_seed(6)
    data = dict(imgs=[torch.rand(4, 3, 8, 32, 32, generator=g) for _ in range(2)],
                flow_imgs=[torch.rand(4, 3, 8, 32, 32, generator=g) for _ in range(2)])
    for step in range(2):
        torch.manual_seed(100 + step)
        out = ref_model.train_step(data, None)
        torch.manual_seed(100 + step)
        loss, log_vars = orc.train_step(data["imgs"][0], data["imgs"][1], data["flow_imgs"][0], data["flow_imgs"][1])
        assert list(log_vars) == list(out["log_vars"])
        for k, v in out["log_vars"].items():
            assert abs(log_vars[k] - v) <= 2e-5 * max(1.0, abs(v)), (kind, step, k, log_vars[k], v)
        for rec, st in ((ref_model.recognizer, orc.rgb.state), (ref_model.recognizer_flow, orc.flow.state)):
            assert int(rec.queue_ptr) == st.ptr and rec.iters == st.iters and rec.batch_size == st.batch_size
            np.testing.assert_array_equal(rec.count.numpy(), st.count.numpy())
            np.testing.assert_allclose(rec.queue.numpy(), st.queue.numpy(), rtol=0, atol=2e-6)


def test_r50_config_full_step_matches_reference():
    """`mscl_r50_cosm_lr3e-2.py`'s model dict (SlowOnly-R50 + TPN with one pyramid convolution, r2d_50 flow branch, LMCL
    head with a Conv1d(256,128) flow projection) built by the unmodified reference and by this repo: identical
    state_dict keys, and the oracle's step (oracle/step.py, host copies of this repo's modules) reproduces the
    reference's train_step at a reduced clip size."""
    import mscl_b200
    from mscl_b200.configs import mscl_r50_model
    from oracle.step import OracleMSCL
    ref = ref_shim.load_reference()
    ref_shim.load_slowonly()                       # registers ResNet3dSlowOnly into the reference's BACKBONES
    ref_shim.ensure_process_group()
    cfg = mscl_r50_model(K=64, aug="IdentityAug")
    cfg["recognizer"]["max_iters"] = cfg["recognizer_flow"]["max_iters"] = 1000
    torch.manual_seed(0)
    ref_model = ref.builder.build_model(dict(cfg, aug=dict(type="SyncMoCoAugmentV5")))
    mine = mscl_b200.build_model(cfg)
    sd_ref = ref_model.state_dict()
    assert set(sd_ref) == set(mine.state_dict())
    mine.load_state_dict(sd_ref, strict=True)
    ref_model.train(), mine.train()
    orc = OracleMSCL(mine)
    g = torch.Generator().manual_seed(8)
    N = 2
    data = dict(imgs=[torch.rand(N, 3, 8, 64, 64, generator=g) for _ in range(2)],
                flow_imgs=[torch.rand(N, 3, 16, 64, 64, generator=g) for _ in range(2)])
    torch.manual_seed(100)
    out = ref_model.train_step(data, None)
    out["loss"].backward()
    torch.manual_seed(100)
    loss, log_vars = orc.train_step(data["imgs"][0], data["imgs"][1], data["flow_imgs"][0], data["flow_imgs"][1])
    loss.backward()
    assert list(log_vars) == list(out["log_vars"]) and len(log_vars) == 23
    for k, v in out["log_vars"].items():
        assert abs(log_vars[k] - v) <= 5e-5 * max(1.0, abs(v)), (k, log_vars[k], v)
    for rec, st in ((ref_model.recognizer, orc.rgb.state), (ref_model.recognizer_flow, orc.flow.state)):
        assert int(rec.queue_ptr) == st.ptr and rec.iters == st.iters and rec.batch_size == st.batch_size
        np.testing.assert_array_equal(rec.count.numpy(), st.count.numpy())
    gr = dict(ref_model.named_parameters())
    a, b = orc.trans_flow.weight.grad, gr["sup_head.trans_flow.weight"].grad
    assert float((a - b).norm() / b.norm()) < 2e-4


def test_moco_head_v2_materialised_form_matches_reference():
    """MoCoHeadV2 (heads/moco_head_v3.py:15-85; its line 8 imports a package that does not exist and is skipped): the
    materialised `forward(q, k, weight)` + `loss(cls_score, ssl_label)` of this repo's class against the reference's,
    on host tensors (plain PyTorch on both sides; the fused path is checked on the GPU against this same form)."""
    import sys
    import mscl_b200
    ref = ref_shim.load_reference()
    if "mmaction.models.heads.moco_head_v3" not in sys.modules:
        ref_shim._load("mmaction.models.heads.moco_head_v3", "mmaction/models/heads/moco_head_v3.py", strip_lines=(8,))
    RefHead = sys.modules["mmaction.models.heads.moco_head_v3"].MoCoHeadV2
    loss_cfg = dict(type="CrossEntropyLoss_torch", ignore_index=-1)
    rh = RefHead(basename="v2", loss_cls=loss_cfg, T=0.2)
    mh = mscl_b200.build_head(dict(type="MoCoHeadV2", basename="v2", loss_cls=loss_cfg, T=0.2))
    g = torch.Generator().manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(16, 128, generator=g), dim=1)
    k = torch.nn.functional.normalize(q + 0.5 * torch.randn(16, 128, generator=g), dim=1)
    w = torch.nn.functional.normalize(torch.randn(128, 512, generator=g), dim=0) * 0.9
    qa, qb = q.clone().requires_grad_(True), q.clone().requires_grad_(True)
    ra, rb = rh(qa, k, w), mh(qb, k, w)
    assert torch.equal(ra["ssl_label"], rb["ssl_label"])
    np.testing.assert_allclose(rb["cls_score"].detach().numpy(), ra["cls_score"].detach().numpy(), rtol=1e-6, atol=1e-6)
    la, lb = rh.loss(**ra), mh.loss(**rb)
    assert list(la.keys()) == list(lb.keys()) == ["top1_acc_v2", "top5_acc_v2", "loss_cls_v2"]
    for key in la:
        assert abs(float(la[key]) - float(lb[key])) <= 1e-6 * max(1.0, abs(float(la[key]))), key
    la["loss_cls_v2"].backward(), lb["loss_cls_v2"].backward()
    np.testing.assert_allclose(qb.grad.numpy(), qa.grad.numpy(), rtol=1e-5, atol=1e-7)


def test_materialised_head_forms_match_reference(pair):
    """The reference-signature paths the fused recognizers bypass but callers holding logits may still use:
    MoCoHead.loss(cls_score, labels), MSCLWithAugMxHead._forward_moco_mx / .loss, and the on-device top-k
    (`#scores above the label < k`) against the reference's NumPy argsort top_k_accuracy."""
    import mscl_b200
    from mscl_b200.heads.moco_head import topk_hits_on_device
    ref = pair[0]
    loss_cfg = dict(type="CrossEntropyLoss_torch", ignore_index=-1)
    g = torch.Generator().manual_seed(1)
    q, k, qf, kf = (torch.nn.functional.normalize(torch.randn(12, 128, generator=g), dim=1) for _ in range(4))
    w, wf = (torch.nn.functional.normalize(torch.randn(128, 300, generator=g), dim=0) for _ in range(2))
    for same_kn in (True, False):
        rh = ref.MSCLWithAugMxHead(basename="mx", loss_cls=loss_cfg, same_kn=same_kn, T=0.07)
        mh = mscl_b200.build_head(dict(type="MSCLWithAugMxHead", basename="mx", loss_cls=loss_cfg, same_kn=same_kn, T=0.07))
        ra, rb = rh._forward_moco_mx(q, k, qf, kf, w, wf), mh._forward_moco_mx(q, k, qf, kf, w, wf)
        for a, b in zip(ra, rb):
            np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-6, atol=1e-6)
        la, lb = rh.loss(*ra, suffix="_aug"), mh.loss(*rb, suffix="_aug")
        assert list(la.keys()) == list(lb.keys())
        for key in la:
            assert abs(float(la[key]) - float(lb[key])) <= 1e-6 * max(1.0, abs(float(la[key]))), (same_kn, key)
    logits = torch.randn(64, 257, generator=g)
    labels = torch.randint(0, 257, (64,), generator=g)
    rhd = ref.MoCoHead(basename="x", loss_cls=loss_cfg)
    mhd = mscl_b200.build_head(dict(type="MoCoHead", basename="x", loss_cls=loss_cfg))
    la, lb = rhd.loss(logits, labels), mhd.loss(logits, labels)
    assert list(la.keys()) == list(lb.keys())
    for key in la:
        assert abs(float(la[key]) - float(lb[key])) <= 1e-6 * max(1.0, abs(float(la[key]))), key
    want = ref.top_k_accuracy(logits.numpy(), labels.numpy(), (1, 5))
    got = [float(v) for v in topk_hits_on_device(logits, labels)]
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-7)
