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
