"""The acceptance config builds unchanged through the product's registry (CPU: construction only)."""
import os

import pytest
import torch

import mscl_b200
from mscl_b200.configs import mscl_r18_model

REF_CFG = "/root/reference/configs/recognition/moco/mscl_r18_cosm_lr2e-2.py"


def _plain(x):
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    return x


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not present")
def test_reference_config_file_loads_and_equals_builder():
    cfg = mscl_b200.Config.fromfile(REF_CFG)
    assert cfg.model.type == "MSCLWithAug" and cfg.data.videos_per_gpu == 32
    assert cfg.dist_params["backend"] == "nccl"            # inherited from _base_/default_runtime.py
    assert cfg.get("find_unused_parameters") is True
    assert _plain(cfg.model) == _plain(mscl_r18_model())
    cfg.merge_from_dict({"model.recognizer.K": 1024, "total_epochs": 1})
    assert cfg.model.recognizer.K == 1024 and cfg.total_epochs == 1
    assert "MSCLWithAug" in cfg.pretty_text


def test_model_builds_with_reference_names():
    model = mscl_b200.build_model(mscl_r18_model(K=512))
    assert type(model).__name__ == "MSCLWithAug"
    assert type(model.recognizer).__name__ == "MoCoV2" and model.recognizer.K == 512
    sd = model.state_dict()
    for key, shape, dtype in (("recognizer.queue", (128, 512), torch.float32), ("recognizer.queue_ptr", (1,), torch.int64),
                              ("recognizer.count", (512,), torch.int64), ("recognizer_flow.queue", (128, 512), torch.float32),
                              ("sup_head.labels", (1, 4), torch.int64)):
        assert tuple(sd[key].shape) == shape and sd[key].dtype == dtype, key
    assert any(k.startswith("recognizer.encoder_q.layer1.0.conv1.0") for k in sd)
    assert not any(p.requires_grad for p in model.recognizer.encoder_k.parameters())
    assert sum(p.numel() for p in model.parameters()) == 74_885_024
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 37_442_512
    # queue columns start unit-norm with age zero (moco.py:390-396)
    assert torch.allclose(sd["recognizer.queue"].norm(dim=0), torch.ones(512), atol=1e-5)
    assert int(sd["recognizer.count"].abs().sum()) == 0 and int(sd["recognizer.queue_ptr"]) == 0


def test_product_refuses_cpu():
    from mscl_b200 import functional as fx, _cabi
    with pytest.raises(_cabi.MsclError):
        fx.hw_mean(torch.zeros(2, 4, 2, 3, 3))
    model = mscl_b200.build_model(mscl_r18_model(K=256, aug="IdentityAug"))
    with pytest.raises(_cabi.MsclError):
        model.recognizer.negative_queue()


REF_CFG_DIR = "/root/reference/configs/recognition/moco"
ALL_CFGS = ["mscl_r18_cosm_lr2e-2.py", "mscl_r50_cosm_lr3e-2.py", "moco_r18_lr3e-2.py",
            "moco_r18_consistent_augmentation_lr3e-2.py", "moco_r18_cosistent_video_lr3e-2.py",
            "moco_r50_consistent_augmentation_lr3e-2.py"]


@pytest.mark.skipif(not os.path.isdir(REF_CFG_DIR), reason="reference tree not present")
@pytest.mark.parametrize("name", ALL_CFGS)
def test_every_moco_config_of_the_reference_builds(name):
    """All six files of configs/recognition/moco load unchanged and build through the product's registry."""
    cfg = mscl_b200.Config.fromfile(os.path.join(REF_CFG_DIR, name))
    keys = ["model.K"] if cfg.model.type == "MoCo" else ["model.recognizer.K", "model.recognizer_flow.K"]
    cfg.merge_from_dict({k: 512 for k in keys})              # smaller queues: construction only
    model = mscl_b200.build_model(cfg.model)
    assert type(model).__name__ == cfg.model.type
    recs = [model] if cfg.model.type == "MoCo" else [model.recognizer, model.recognizer_flow]
    for rec in recs:
        assert not any(p.requires_grad for m in (rec.encoder_k, rec.neck_k, rec.mlp_k) for p in m.parameters())
        for pq, pk in zip(rec.encoder_q.parameters(), rec.encoder_k.parameters()):
            assert torch.equal(pq, pk)
    if cfg.model.type == "MoCo":
        assert model.m == 0.999 and type(model.aug_gpu).__name__ == cfg.model.aug.type
    if name == "mscl_r50_cosm_lr3e-2.py":
        # BASELINE config 5's parameter list: SlowOnly-R50 31,672,128 elements / 159 tensors + TPN neck + MLP
        r = model.recognizer
        assert sum(p.numel() for p in r.encoder_k.parameters()) == 31_672_128
        assert len(list(r.encoder_k.parameters())) == 159
        total = sum(p.numel() for m in (r.encoder_k, r.neck_k, r.mlp_k) for p in m.parameters())
        assert 38_000_000 < total < 39_000_000, total
        assert _plain(cfg.model.recognizer.backbone) == _plain(mscl_b200.configs.mscl_r50_model()["recognizer"]["backbone"])


def test_sibling_registry_names():
    """the sibling heads / recognizers / augmentations resolve under the reference's names."""
    for n in ("MoCo", "MoCoV2", "MSCL", "MSCLWithAug", "MoDist", "MoCoHead", "MoCoHeadV2", "MSCLWithAugMxHead",
              "MSCLWithAugPosHeadV2", "MSCLWithAugPosHead", "MoDistv2PosHead", "MlvlMSCLWithAugPosHead",
              "MSCLWithAugSimpleHead", "ResNet3dSlowOnly", "TPNMoCo", "BaseMoCo", "CrossEntropyLoss_torch"):
        assert n in mscl_b200.MODELS, n
    for n in ("IdentityAug", "SyncMoCoAugmentV5", "SyncMoCoAugmentV2", "MoCoAugmentV2"):
        assert n in mscl_b200.SSL_AUGS, n


def test_augmentations_refuse_host_tensors():
    from mscl_b200 import _cabi
    x = torch.rand(2, 3, 4, 16, 16)
    for cfg in (dict(type="SyncMoCoAugmentV5", crop_size=16, t=(4, 4)), dict(type="SyncMoCoAugmentV2", crop_size=16, t=4),
                dict(type="MoCoAugmentV2", crop_size=16)):
        aug = mscl_b200.build_ssl_aug(cfg)
        with pytest.raises(_cabi.MsclError):
            aug(x, x, {})


def test_augmentation_host_draw_and_pack():
    """SyncMoCoAugmentV5 draws its decisions on the host and uploads one packed table per view: the host packing
    (`_pack_host`) equals the device-op packing the kernel tests use (`_pack_params`), the draws respect the reference's
    ranges and sharing rules (ssl_aug_v2.py:31-48: ColorJitter(0.4,0.4,0.4,0.1) p=.8, grayscale p=.2, blur p=.5; 'batch'
    level: decisions shared by the frames of a clip, factors per frame), and `torch.manual_seed` makes them repeatable
    without touching the CPU default generator (which draws the shuffle permutations)."""
    import numpy as np
    from mscl_b200.common.ssl_aug import SyncMoCoAugmentV5
    n, t = 64, 8
    draws = []
    for _ in range(2):
        torch.manual_seed(7)
        aug = SyncMoCoAugmentV5(crop_size=112, sync_level=("batch", "params"), t=(8, 8), flow_suffix="flow_imgs")
        before = torch.get_rng_state()
        d = aug._draw_host(n, t)
        assert torch.equal(before, torch.get_rng_state())
        draws.append(d)
    for k in draws[0]:
        assert np.array_equal(np.asarray(draws[0][k]), np.asarray(draws[1][k])), k
    d = draws[0]
    for k in ("brightness", "contrast", "saturation"):
        assert d[k].shape == (n * t,) and d[k].min() >= 0.6 and d[k].max() <= 1.4
        assert d[k].reshape(n, t).std(axis=1).min() > 0                       # per frame in the 'batch' level
    assert np.abs(d["hue"]).max() <= 0.1 and 0.1 <= d["sigma"] <= 2.0
    for k, p in (("jit", 0.8), ("gray", 0.2), ("blur", 0.5)):
        per_clip = d[k].reshape(n, t)
        assert (per_clip == per_clip[:, :1]).all()                            # decisions shared by the frames of a clip
        assert abs(per_clip[:, 0].mean() - p) < 0.2
    assert d["taps"].shape == (11,) and abs(float(d["taps"].sum()) - 1) < 1e-6 and np.allclose(d["taps"], d["taps"][::-1])
    assert aug._draw_host(n, 1)["brightness"].shape == (n,)                   # 'params' level: one set per clip
    flip = np.repeat(np.arange(n) % 3 == 0, t)
    rows = aug._pack_host(d, flip, False)
    prm = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in d.items()}
    ref = aug._pack_params(prm, torch.from_numpy(flip), False).numpy()
    np.testing.assert_allclose(rows, ref, rtol=0, atol=1e-6)
    weak = aug._pack_host(d, flip, True)
    assert np.array_equal(weak[:, 0], flip.astype(np.float32)) and not weak[:, 1:].any()
    prm2 = aug._color_params(4, torch.device("cpu"), 8)
    assert prm2["jit"].dtype == torch.bool and prm2["brightness"].shape == (32,) and prm2["taps"].shape == (11,)
