"""MSCLWithAug: RGB MoCoV2 + flow MoCoV2 (base flow and FRA-rotated flow) + cross-modal
InfoNCE + LMCL (reference: recognizers/mscl.py:137-292), and MSCL, the same model without the
FRA branch (recognizers/mscl.py:9-134).

The reference evaluates 7 InfoNCE terms per step, each with its own materialised
(N,1+K) logits, against only THREE distinct negative matrices:

    W_rgb  (before this step's RGB enqueue)   <- q (own), q_f (fr), q_af (fr_aug)
    W_flow (before the base-flow enqueue)     <- q_f (own)
    W_flow (after the base-flow enqueue)      <- q_af (own_aug), q (rf), q (rf_aug)

Here each matrix is streamed ONCE by the fused kernel with the query sets stacked
(3N + N + 3N rows); the RGB enqueue is merely deferred until the rows that need the
pre-enqueue RGB queue exist.  Effects visible from outside (queue contents, pointer, ages,
`iters`, momentum, permutation draws, loss keys and values) are those of the reference.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist

from ..registry import RECOGNIZERS, build_ssl_aug
from .base_moco import TwoBranchRecognizer


def two_branch_rows(rec, recf, q, k, q_f, k_f, same_kn, T_mx):
    """The four InfoNCE terms of a two-branch step without the FRA call (mscl.py:92-113, modist.py:84-118): own RGB,
    own flow, rf (q vs k_flow) and fr (q_flow vs k).  Every term reads a PRE-enqueue decayed queue (each recognizer
    snapshots its weight before its own enqueue, moco.py:484-502), so each queue is streamed once with the two row
    sets that need it stacked; then both enqueues run.  Returns {name: [loss, top1, top5, 0] row}."""
    for head in (rec.moco_head, recf.moco_head):
        if not head.can_fuse():
            raise NotImplementedError("the fused path needs loss_cls=CrossEntropyLoss_torch without class weights")
    rf_queue, fr_queue = ("flow", "rgb") if same_kn else ("rgb", "flow")
    terms = {"rgb": [("own", q, k, rec.T)], "flow": [("own_f", q_f, k_f, recf.T)]}
    terms[rf_queue].append(("rf", q, k_f, T_mx))
    terms[fr_queue].append(("fr", q_f, k, T_mx))
    rows = {}
    for phase, owner in (("rgb", rec), ("flow", recf)):
        by_T = OrderedDict()          # one pass per distinct temperature (one, in the configs)
        for name, qq, kk, T in terms[phase]:
            by_T.setdefault(T, []).append((name, qq, kk))
        for T, items in by_T.items():
            out = owner.contrast([(qq, kk) for _, qq, kk in items], T)
            for i, item in enumerate(items):
                rows[item[0]] = out[i]
    rec._dequeue_and_enqueue(k)
    recf._dequeue_and_enqueue(k_f)
    return rows


def _wants(head, name):
    """Does the head's aux_keys map ask for feature `name` (e.g. the unshuffled key pyramid `k_mlvl`)?"""
    return any(name in v for v in getattr(head, "aux_keys", {}).values())


@RECOGNIZERS.register_module()
class MSCL(TwoBranchRecognizer):
    """RGB MoCo + flow MoCo + cross-modal InfoNCE (`moco_mx_head`) + a frame-level head (`sup_head`, e.g.
    MoDistv2PosHead) on the base flow only (recognizers/mscl.py:9-134)."""

    def __init__(self, recognizer, recognizer_flow, moco_mx_head, sup_head, im_key="imgs", flow_key="flows",
                 flow_img_key="flow_imgs", aux_info=[], aug=dict(dtype="MoCoAugmentV3", moco_aug=(112, 112), t=8),
                 same_kn=True, update_aug_flow=False, weight_aug_flow=(1.0, 1.0), train_cfg=None, test_cfg=None):
        super().__init__(train_cfg=train_cfg, test_cfg=test_cfg)
        self._build_branches(recognizer, recognizer_flow, train_cfg)
        self.im_key = im_key
        self.same_kn = same_kn
        self.update_aug_flow = update_aug_flow        # stored, unused (as the reference)
        self.weight_aug_flow = weight_aug_flow
        self.flow_key = flow_key
        self.flow_img_key = flow_img_key
        self.aux_info = aux_info
        self._build_cls_head(moco_mx_head, name="moco_mx_head")
        self._build_cls_head(sup_head, name="sup_head")
        self.aug_gpu = build_ssl_aug(aug)

    def train_step(self, data_batch, optimizer, **kwargs):
        im_q = data_batch[self.im_key][0]
        im_k = data_batch[self.im_key][1]
        aux_info = {f"{self.flow_key}_q": data_batch[self.flow_key][0], f"{self.flow_key}_k": data_batch[self.flow_key][1]}
        aux_info.update(self._collect_aux(data_batch))
        return self._finish_step(self(im_q, im_k, aux_info, return_loss=True), im_q.shape[0])

    def objective(self, feats):
        """Everything after the encoders (mscl.py:92-120).  feats: q, k, q_f, k_f (N,128) and the feature dicts
        `im_features` / `flow_features` handed to the frame-level head."""
        rec, recf, mx = self.recognizer, self.recognizer_flow, self.moco_mx_head
        if not mx.can_fuse():
            raise NotImplementedError("the fused path needs loss_cls=CrossEntropyLoss_torch without class weights")
        rows = two_branch_rows(rec, recf, feats["q"], feats["k"], feats["q_f"], feats["k_f"], mx.same_kn, mx.T)
        losses = OrderedDict()
        losses.update(rec.moco_head.loss_fused(rows["own"]))
        losses.update(recf.moco_head.loss_fused(rows["own_f"]))
        losses.update(mx.loss_fused_mx(rows["rf"], rows["fr"]))
        aux = dict(feats.get("aux_info") or {})
        aux = self.sup_head.update_aux_info("im_features", feats["im_features"], aux)
        aux = self.sup_head.update_aux_info("base_flow_features", feats["flow_features"], aux)
        aux.update(self.sup_head(**aux))
        losses.update(self.sup_head.loss(**aux))
        return losses

    def forward_train(self, im_q, im_k, aux_info):
        im_q, im_k, aux_info = self.aug_gpu(im_q, im_k, aux_info)
        rec, recf = self.recognizer, self.recognizer_flow
        flow_q, flow_k = aux_info[f"{self.flow_img_key}_q"], aux_info[f"{self.flow_img_key}_k"]
        n = im_q.shape[0]
        need_k = _wants(self.sup_head, "k_mlvl")
        q, q_mlvl, k, k_mlvl, _ = rec.extract_feat(im_q, im_k, unshuffle_mlvl=need_k)
        rec.note_branch(n, True)
        q_f, qf_mlvl, k_f, kf_mlvl, _ = recf.extract_feat(flow_q, flow_k, unshuffle_mlvl=need_k)
        recf.note_branch(n, True)
        return self.objective(dict(q=q, k=k, q_f=q_f, k_f=k_f, aux_info=aux_info,
                                   im_features=dict(q=q, q_mlvl=q_mlvl, k=k, k_mlvl=k_mlvl, q_neg=None),
                                   flow_features=dict(q=q_f, q_mlvl=qf_mlvl, k=k_f, k_mlvl=kf_mlvl, q_neg=None)))



@RECOGNIZERS.register_module()
class MSCLWithAug(TwoBranchRecognizer):
    def __init__(self, recognizer, recognizer_flow, moco_mx_head, sup_head, im_key="imgs", flow_key="flow_imgs",
                 aux_info=[], aug=dict(dtype="MoCoAugmentV3", moco_aug=(112, 112), t=8), same_kn=True,
                 update_aug_flow=False, weight_aug_flow=(1.0, 1.0), train_cfg=None, test_cfg=None):
        super().__init__(train_cfg=train_cfg, test_cfg=test_cfg)
        self._build_branches(recognizer, recognizer_flow, train_cfg)       # options such as shard_queue reach both
        self.im_key = im_key
        self.same_kn = same_kn
        self.update_aug_flow = update_aug_flow
        self.weight_aug_flow = weight_aug_flow
        if isinstance(flow_key, (list, tuple)):
            self.cat_flow = False
            self.flow_key = flow_key
        else:
            self.cat_flow = True
            self.flow_key = (flow_key,)
        self.aux_info = aux_info
        self._build_cls_head(moco_mx_head, name="moco_mx_head")
        self._build_cls_head(sup_head, name="sup_head")
        self.aug_gpu = build_ssl_aug(aug)
        # train_cfg=dict(merge_flow_epochs=False) keeps the three-pass schedule (W_flow streamed before AND after its enqueue)
        self.merge_flow_epochs = bool((train_cfg or {}).get("merge_flow_epochs", True))

    def train_step(self, data_batch, optimizer, **kwargs):
        im_q = data_batch[self.im_key][0]
        im_k = data_batch[self.im_key][1]
        aux_info = {}
        for flow_key in self.flow_key:
            aux_info[f"{flow_key}_q"] = data_batch[flow_key][0]
            aux_info[f"{flow_key}_k"] = data_batch[flow_key][1]
        aux_info.update(self._collect_aux(data_batch))
        return self._finish_step(self(im_q, im_k, aux_info, return_loss=True), im_q.shape[0])

    def objective(self, feats):
        """Everything after the encoders (mscl.py:228-277 + moco.py:481-510), fused.

        feats: q,k (RGB) q_f,k_f (base flow) q_af,k_af (FRA flow), each (N,128), and the
        multi-level query features q_mlvl, q_flow_mlvl, q_aug_flow_mlvl.
        """
        rec, recf, mx = self.recognizer, self.recognizer_flow, self.moco_mx_head
        q, k, q_f, k_f, q_af, k_af = (feats[n] for n in ("q", "k", "q_f", "k_f", "q_af", "k_af"))
        for head in (rec.moco_head, recf.moco_head, mx):
            if not head.can_fuse():
                raise NotImplementedError("the fused path needs loss_cls=CrossEntropyLoss_torch without class weights")
        use_aug_mx = self.weight_aug_flow[1] > 0

        # which decayed queue each cross-modal term reads (heads/moco_head_v2.py:42-47)
        # ... decided by the HEAD's same_kn, as in the reference (the recognizer's own `same_kn` argument is stored, unused)
        rf_queue, fr_queue = ("flow_post", "rgb_pre") if mx.same_kn else ("rgb_pre", "flow_post")
        terms = {"rgb_pre": [("own", q, k, rec.T)], "flow_pre": [("own_f", q_f, k_f, recf.T)],
                 "flow_post": [("own_af", q_af, k_af, recf.T)]}
        terms[rf_queue].append(("rf", q, k_f, mx.T))
        terms[fr_queue].append(("fr", q_f, k, mx.T))
        if use_aug_mx:
            terms[rf_queue].append(("rf_aug", q, k_af, mx.T))
            terms[fr_queue].append(("fr_aug", q_af, k, mx.T))

        rows = {}
        # A term evaluated against W_flow(post) whose positive key IS k_f finds a copy of that key in
        # the queue (`rf` with same_kn=True): the kernel needs the slot to rank it exactly.
        kf_slots = recf.enqueue_slots(k_f.shape[0], k_f.device)

        def passes(phase):
            by_T = OrderedDict()          # one pass per distinct temperature (one, in the configs)
            for name, qq, kk, T in terms[phase]:
                dup = kf_slots if (phase == "flow_post" and kk is k_f) else None
                by_T.setdefault(T, []).append((name, qq, kk, dup))
            return list(by_T.items())

        def launch(calls, names):
            for out, nm in zip(type(rec).contrast_many(calls), names):
                for i, name in enumerate(nm):
                    rows[name] = out[i]

        def run(*phase_owner):
            """The passes of the given (phase, recognizer) pairs; passes over different queues that do not depend on
            each other go out as ONE launch (MoCoV2.contrast_many)."""
            calls, names = [], []
            for phase, owner in phase_owner:
                for T, items in passes(phase):
                    calls.append((owner, [(qq, kk, dup) for _, qq, kk, dup in items], T))
                    names.append([item[0] for item in items])
            launch(calls, names)

        pre_f, post_f = passes("flow_pre"), passes("flow_post")
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        if (self.merge_flow_epochs and len(passes("rgb_pre")) == 1 and len(pre_f) == 1 and len(post_f) == 1 and pre_f[0][0] == post_f[0][0]
                and k_f.shape[0] * world <= 128 and type(rec).can_launch_together([rec, recf], q.device)):
            # W_flow before and after the base-flow enqueue differ in B slots and one unit of age: ONE pass over it serves
            # both row sets (mscl_infonce_fused_multi_x) -- the enqueue runs first and hands over what it overwrote -- in
            # the same launch as the W_rgb pass: the whole step's InfoNCE is one launch over two queues instead of three
            # passes.  (The RGB enqueue still follows its consumers.)
            overwritten = recf._dequeue_and_enqueue(k_f, save=True)      # base-flow call's enqueue
            calls = [(rec, [(qq, kk, dup) for _, qq, kk, dup in items], T) for T, items in passes("rgb_pre")]
            names = [[item[0] for item in items] for _, items in passes("rgb_pre")]
            both = pre_f[0][1] + post_f[0][1]
            calls.append((recf, [(qq, kk, dup) for _, qq, kk, dup in both], pre_f[0][0], (overwritten, len(pre_f[0][1]))))
            names.append([item[0] for item in both])
            launch(calls, names)
            rec._dequeue_and_enqueue(k)                # RGB call's enqueue (deferred past its consumers)
        else:
            # W_rgb before this step's enqueue and W_flow before the base-flow enqueue: two independent passes, one launch
            run(("rgb_pre", rec), ("flow_pre", recf))
            rec._dequeue_and_enqueue(k)                # RGB call's enqueue (deferred past its consumers)
            recf._dequeue_and_enqueue(k_f)             # base-flow call's enqueue
            run(("flow_post", recf))                   # W_flow containing this step's base-flow keys
        if self.update_aug_flow:
            recf._dequeue_and_enqueue(k_af)

        losses = OrderedDict()
        losses.update(rec.moco_head.loss_fused(rows["own"]))
        loss_flow = recf.moco_head.loss_fused(rows["own_f"])
        for key, val in recf.moco_head.loss_fused(rows["own_af"]).items():
            if key.startswith("loss"):          # the FRA branch's accuracies are dropped (mscl.py:242-245)
                loss_flow[key + "_aug"] = val * self.weight_aug_flow[0]
        losses.update(loss_flow)
        losses.update(mx.loss_fused_mx(rows["rf"], rows["fr"]))
        if use_aug_mx:
            losses.update(mx.loss_fused_mx(rows["rf_aug"], rows["fr_aug"], suffix="_aug"))

        aux = {}
        aux = self.sup_head.update_aux_info("im_features", dict(q_mlvl=feats["q_mlvl"]), aux)
        aux = self.sup_head.update_aux_info("base_flow_features", dict(q_mlvl=feats["q_flow_mlvl"]), aux)
        aux = self.sup_head.update_aux_info("aug_flow_features", dict(q_mlvl=feats["q_aug_flow_mlvl"]), aux)
        aux.update(self.sup_head(**aux))
        losses.update(self.sup_head.loss(**aux))
        return losses

    def forward_train(self, im_q, im_k, aux_info):
        im_q, im_k, aux_info = self.aug_gpu(im_q, im_k, aux_info)
        rec, recf = self.recognizer, self.recognizer_flow
        if self.cat_flow:
            cat_q, cat_k = aux_info[f"{self.flow_key[0]}_q"], aux_info[f"{self.flow_key[0]}_k"]
            # strided views: the encoder's first convolution (or the graphed path's static-input copy) makes the one
            # dense copy it needs in its own layout; a `.contiguous()` here would be a second 38 MB pass per half
            flow_q, aug_flow_q = cat_q.chunk(2, 2)
            flow_k, aug_flow_k = cat_k.chunk(2, 2)
        else:
            flow_q, flow_k = aux_info[f"{self.flow_key[0]}_q"], aux_info[f"{self.flow_key[0]}_k"]
            aug_flow_q, aug_flow_k = aux_info[f"{self.flow_key[1]}_q"], aux_info[f"{self.flow_key[1]}_k"]
        # encoders in the reference's order: each call updates its key encoder by EMA and draws
        # one shuffle permutation (RGB, base flow, FRA flow); `iters` advances after each call so the
        # second flow EMA sees the advanced schedule
        n = im_q.shape[0]
        q, q_mlvl, k, _, _ = rec.extract_feat(im_q, im_k, unshuffle_mlvl=False)
        rec.note_branch(n, True)
        q_f, qf_mlvl, k_f, _, _ = recf.extract_feat(flow_q, flow_k, unshuffle_mlvl=False)
        recf.note_branch(n, True)
        q_af, qaf_mlvl, k_af, _, _ = recf.extract_feat(aug_flow_q, aug_flow_k, unshuffle_mlvl=False, site=1)
        recf.note_branch(n, self.update_aug_flow)
        return self.objective(dict(q=q, k=k, q_f=q_f, k_f=k_f, q_af=q_af, k_af=k_af, q_mlvl=q_mlvl,
                                   q_flow_mlvl=qf_mlvl, q_aug_flow_mlvl=qaf_mlvl))

