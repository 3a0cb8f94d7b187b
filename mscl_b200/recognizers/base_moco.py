"""Base class of the MoCo-style recognizers (recognizers/base_moco.py:10-162 and the
`_parse_losses` of recognizers/base.py:275-308).

Same builder helpers and backbone-prefix resolution as the reference.  `_parse_losses`
returns the same (loss, log_vars) but reduces all log variables with ONE all_reduce and ONE
device->host copy instead of one collective + `.item()` per variable (23 per step in the
MSCL config).
"""
import warnings
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn

from .. import registry as builder
from ..backbones import ResNetFlow, torchvision_multilevel


class DeferredLogVars:
    """The step's log variables as one stacked tensor.  Their copy to pinned host memory is enqueued at once, followed by an
    event; `.get()` waits for THAT event only.  (`stacked.tolist()` at `.get()` time would be a copy ordered behind
    everything enqueued since -- a caller that reads them one step late, to keep the host a step ahead of the device, would
    wait for the whole next step instead: the GPU then idled through the eager start of every step,
    profiles/r02_timeline_host_g1.txt.)"""

    _ring, _ring_pos = [], 0          # pinned landing buffers + the event of the copy that last wrote each (16 slots)

    def __init__(self, keys, stacked):
        self.keys, self.stacked, self._values = keys, stacked, None
        self._host = self._event = None
        if stacked.is_cuda:
            cls = DeferredLogVars
            if not cls._ring:
                cls._ring = [[torch.empty(256, dtype=torch.float32, pin_memory=True), None, 0] for _ in range(16)]
            slot = cls._ring[cls._ring_pos]
            cls._ring_pos = (cls._ring_pos + 1) % len(cls._ring)
            if stacked.numel() > slot[0].numel() or stacked.dtype != torch.float32:
                slot = [torch.empty(stacked.shape, dtype=stacked.dtype, pin_memory=True), None, 0]   # unusual: own buffer
            elif slot[1] is not None:
                slot[1].synchronize()          # 16 steps old: long done
            slot[2] += 1                       # whoever still holds the slot's previous contents falls back to `stacked`
            self._slot, self._gen = slot, slot[2]
            self._host = slot[0][:stacked.numel()].view(stacked.shape)
            self._host.copy_(stacked, non_blocking=True)
            self._event = slot[1] = torch.cuda.Event()
            self._event.record()

    def get(self):
        if self._values is None:
            if self._event is not None and self._slot[2] == self._gen:
                self._event.synchronize()
                values = self._host.tolist()
            else:
                values = self.stacked.tolist()
            self._values = OrderedDict(zip(self.keys, values))
        return self._values


class BaseMoCoRecognizer(nn.Module):
    def __init__(self, backbone=None, cls_head=None, neck=None, train_cfg=None, test_cfg=None):
        super().__init__()
        self.backbone_from = "mmaction2"
        self.backbone_list, self.neck_list, self.cls_head_list = [], [], []
        if backbone is not None:
            self._build_backbone(backbone)
        if neck is not None:
            self._build_neck(neck)
        if cls_head is not None:
            self._build_cls_head(cls_head)
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.aux_info = []
        if train_cfg is not None and "aux_info" in train_cfg:
            self.aux_info = train_cfg["aux_info"]
        self.fp16_enabled = False

    # -- builders (recognizers/base_moco.py:77-114) --
    def _build_backbone(self, backbone, name="backbone"):
        backbone = dict(backbone)
        typ = backbone["type"]
        if typ.startswith("torchvision."):
            from torchvision.models import video as tv_video
            net = tv_video.__dict__[backbone.pop("type")[12:]](**backbone)
            net.classifier = nn.Identity()
            net.fc = nn.Identity()
            self.backbone_from = "torchvision"
            setattr(self, name, torchvision_multilevel(net))
        elif typ.startswith("resnet_flow."):
            net = ResNetFlow(backbone.pop("type")[12:], **backbone)
            net.classifier = nn.Identity()
            net.fc = nn.Identity()
            self.backbone_from = "torchvision"
            setattr(self, name, net)
        else:
            setattr(self, name, builder.build_backbone(backbone))
        self.backbone_list.append(name)

    def _build_neck(self, neck, name="neck"):
        setattr(self, name, builder.build_neck(neck))
        self.neck_list.append(name)

    def _build_cls_head(self, cls_head, name="cls_head"):
        setattr(self, name, builder.build_head(cls_head))
        self.cls_head_list.append(name)

    @property
    def with_neck(self):
        return len(self.neck_list) > 0

    @property
    def with_cls_head(self):
        return len(self.cls_head_list) > 0

    def init_weights(self):
        for bn in self.backbone_list:
            if self.backbone_from in ("mmcls", "megaction", "mmaction2"):
                getattr(self, bn).init_weights()
            elif self.backbone_from != "torchvision":
                raise NotImplementedError(f"Unsupported backbone source {self.backbone_from}!")
        for n in self.neck_list + self.cls_head_list:
            getattr(self, n).init_weights()

    # -- loss parsing (recognizers/base.py:275-308) --
    @staticmethod
    def parse_losses_deferred(losses):
        """`_parse_losses` without its device->host copy: returns (loss, DeferredLogVars).  The caller may enqueue
        the backward pass and the optimizer before `.get()` reads the step's log variables, so the host never waits
        for the forward pass in the middle of a step (the reference's `.item()` per variable does, base.py:301-306)."""
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f"{name} is not a tensor or list of tensors")
        loss = sum(v for k, v in log_vars.items() if "loss" in k)
        log_vars["loss"] = loss
        stacked = torch.stack([v.detach().float() for v in log_vars.values()])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            stacked = stacked / dist.get_world_size()
            dist.all_reduce(stacked)
        return loss, DeferredLogVars(list(log_vars.keys()), stacked)

    @staticmethod
    def _parse_losses(losses):
        loss, deferred = BaseMoCoRecognizer.parse_losses_deferred(losses)
        return loss, deferred.get()          # the step's single device->host synchronisation

    def forward_train(self, imgs, labels, **kwargs):
        raise NotImplementedError("Not support forward_train for BaseMoCoRecognizer")

    def forward_test(self, imgs):
        raise NotImplementedError("Not support forward_test for BaseMoCoRecognizer")

    def forward_gradcam(self, imgs):
        raise NotImplementedError("Not support forward_gradcam for BaseMoCoRecognizer")


class TwoBranchRecognizer(BaseMoCoRecognizer):
    """What MSCL, MSCLWithAug and MoDist share: an RGB and a flow MoCo recognizer built from config dicts (options in
    `train_cfg`, e.g. shard_queue, reach both), the `forward` dispatch of the SSL recognizers and their refusals
    (recognizers/mscl.py:74-83,122-134; modist.py:66-75,120-132)."""

    def _build_branches(self, recognizer, recognizer_flow, train_cfg):
        if train_cfg:
            recognizer = dict(recognizer, train_cfg=dict(recognizer.get("train_cfg") or {}, **train_cfg))
            recognizer_flow = dict(recognizer_flow, train_cfg=dict(recognizer_flow.get("train_cfg") or {}, **train_cfg))
        self.recognizer = builder.build_recognizer(recognizer)
        self.recognizer_flow = builder.build_recognizer(recognizer_flow)

    def _collect_aux(self, data_batch):
        aux_info = {}
        for item in self.aux_info:
            assert item in data_batch
            aux_info[item] = data_batch[item]
        return aux_info

    def _finish_step(self, losses, n):
        loss, log_vars = self._parse_losses(losses)
        return dict(num_samples=n, loss=loss, log_vars=log_vars)

    def forward(self, *inputs, return_loss=True, **kwargs):
        if kwargs.pop("gradcam", False):
            return self.forward_gradcam(*inputs, **kwargs)
        if return_loss:
            return self.forward_train(*inputs, **kwargs)
        raise NotImplementedError("MoCo doesnt support test mode")

    def forward_test(self, imgs):
        raise NotImplementedError("Not support for ssl recognizer !!!")

    def forward_gradcam(self, *inputs, **kwargs):
        raise NotImplementedError("Not support for ssl recognizer !!!")

    def extract_global_feat(self):
        raise NotImplementedError("Not support for ssl recognizer !!!")

    def extract_feat(self, im_q, im_k):
        pass

    def visualize(self, data_batch):
        pass


def warn_once(msg, _seen=set()):
    if msg not in _seen:
        _seen.add(msg)
        warnings.warn(msg)
