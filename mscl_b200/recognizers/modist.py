"""MoDist: RGB MoCo + flow MoCo + the two cross-modal InfoNCE terms, no frame-level head
(reference: recognizers/modist.py:9-132), on the fused kernels.

The reference materialises four (N,1+K) logits matrices per step (own RGB, own flow, rf, fr), each against a decayed
queue snapshot taken before that recognizer's enqueue.  Here each queue is streamed ONCE with the two row sets that
read it stacked (`mscl.two_branch_rows`), then both enqueues run; loss keys, their order and every state effect
(queue, pointer, ages, iters, momentum, permutation draws) are the reference's.
"""
from collections import OrderedDict

from ..registry import RECOGNIZERS, build_ssl_aug
from .base_moco import TwoBranchRecognizer
from .mscl import two_branch_rows


@RECOGNIZERS.register_module()
class MoDist(TwoBranchRecognizer):
    def __init__(self, recognizer, recognizer_flow, moco_head, im_key="imgs", flow_key="flow_imgs", aux_info=[],
                 aug=dict(dtype="MoCoAugmentV3", moco_aug=(112, 112), t=8), same_kn=True, train_cfg=None, test_cfg=None):
        super().__init__(train_cfg=train_cfg, test_cfg=test_cfg)
        self._build_branches(recognizer, recognizer_flow, train_cfg)
        self.T = self.recognizer.T
        self.im_key = im_key
        self.flow_key = flow_key
        self.same_kn = same_kn
        self.aux_info = aux_info
        # the reversed-direction head is a copy of the config with `_r` appended to its basename (modist.py:42-45)
        moco_head_r = dict(moco_head)
        moco_head_r["basename"] = moco_head["basename"] + "_r"
        self._build_cls_head(moco_head, name="moco_head")
        self._build_cls_head(moco_head_r, name="moco_head_r")
        self.aug_gpu = build_ssl_aug(aug)

    def train_step(self, data_batch, optimizer, **kwargs):
        im_q, im_k = data_batch[self.im_key][0], data_batch[self.im_key][1]
        flow_q, flow_k = data_batch[self.flow_key][0], data_batch[self.flow_key][1]
        losses = self(im_q, im_k, flow_q, flow_k, self._collect_aux(data_batch), return_loss=True)
        return self._finish_step(losses, im_q.shape[0])

    def objective(self, q, k, q_f, k_f):
        """modist.py:84-118 from the encoder outputs on; losses in the reference's order (rf, fr, RGB, flow)."""
        rec, recf = self.recognizer, self.recognizer_flow
        for head in (self.moco_head, self.moco_head_r):
            if not head.can_fuse():
                raise NotImplementedError("the fused path needs loss_cls=CrossEntropyLoss_torch without class weights")
        rows = two_branch_rows(rec, recf, q, k, q_f, k_f, self.same_kn, self.T)
        losses = OrderedDict()
        losses.update(self.moco_head.loss_fused(rows["rf"]))
        losses.update(self.moco_head_r.loss_fused(rows["fr"]))
        losses.update(rec.moco_head.loss_fused(rows["own"]))
        losses.update(recf.moco_head.loss_fused(rows["own_f"]))
        return losses

    def forward_train(self, im_q, im_k, flow_q, flow_k, aux_info):
        im_q, im_k, flow_q, flow_k, aux_info = self.aug_gpu.forward_with_flow(im_q, im_k, flow_q, flow_k, aux_info)
        rec, recf = self.recognizer, self.recognizer_flow
        n = im_q.shape[0]
        q, _, k, _, _ = rec.extract_feat(im_q, im_k, unshuffle_mlvl=False)
        rec.note_branch(n, True)
        q_f, _, k_f, _, _ = recf.extract_feat(flow_q, flow_k, unshuffle_mlvl=False)
        recf.note_branch(n, True)
        return self.objective(q, k, q_f, k_f)

