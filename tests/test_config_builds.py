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
