"""CPU restatement of one full MSCLWithAug training step.  TEST / BENCH INFRASTRUCTURE ONLY.

The encoders, necks and projection MLPs are ordinary PyTorch modules (they stay in PyTorch in
the product too, so the SAME module classes are instantiated here on the CPU and given the
SAME weights); everything the hot path replaces -- momentum EMA, shuffle-BN, decayed-queue
snapshot, logits, cross-entropy, host top-k, enqueue, LMCL -- is evaluated with the
reference's own operation sequence through oracle/mscl_oracle.py
(mmaction/models/recognizers/moco.py:408-547, mscl.py:225-277).

Used (a) by the GPU parity tests as the checker of a whole step, (b) by bench.py as the
CPU baseline (`--impl reference`), (c) by tests/test_oracle_vs_reference.py, which pins it
against the unmodified reference run through oracle/ref_shim.py.
"""
import copy

import torch
import torch.nn.functional as F

from . import mscl_oracle as O


class OracleBranch:
    """One MoCoV2: q/k encoder, neck, MLP on the CPU + reference-layout queue state.
    device: where the copies live.  "cpu" for every checker use; bench.py's `gpu_eager_baseline` passes the GPU to time the
    reference's operation sequence as eager PyTorch on the same B200 (the GPU-vs-GPU bar)."""

    def __init__(self, rec, state=None, device="cpu"):
        dev = torch.device(device)
        self.encoder_q, self.encoder_k = copy.deepcopy(rec.encoder_q).to(dev), copy.deepcopy(rec.encoder_k).to(dev)
        self.neck_q, self.neck_k = copy.deepcopy(rec.neck_q).to(dev), copy.deepcopy(rec.neck_k).to(dev)
        self.mlp_q, self.mlp_k = copy.deepcopy(rec.mlp_q).to(dev), copy.deepcopy(rec.mlp_k).to(dev)
        # torchvision backbones carry the multi-level forward as an instance attribute bound to the
        # ORIGINAL module; rebind it to the copy
        from mscl_b200.backbones import torchvision_multilevel
        for enc in (self.encoder_q, self.encoder_k):
            if "forward" in enc.__dict__:
                del enc.__dict__["forward"]
                torchvision_multilevel(enc)
        # the product's pyramid convolutions up-sample with the K7 kernel; the host copies take the library op
        for neck in (self.neck_q, self.neck_k):
            for mod in neck.modules():
                if hasattr(mod, "upsample"):
                    mod.upsample = lambda x, size: F.interpolate(x, size=size, mode="trilinear")
        st = state or rec._gathered_state()
        self.state = O.QueueState(st["queue"].to(dev).float(), st["count"].to(dev).long(), int(st["ptr"]))
        self.state.iters, self.state.batch_size = rec.iters, rec.batch_size
        self.m_base, self.max_iters, self.T = rec.m_base, rec.max_iters, rec.T
        # the first MoCo recognizer keeps a constant momentum (moco.py:114-124); MoCoV2 anneals it (:413-415)
        self.fixed_m = rec.m if type(rec).__name__ == "MoCo" else None
        self.train(rec.training)

    def train(self, mode=True):
        for m in (self.encoder_q, self.encoder_k, self.neck_q, self.neck_k, self.mlp_q, self.mlp_k):
            m.train(mode)

    def q_params(self):
        return [p for m in (self.encoder_q, self.neck_q, self.mlp_q) for p in m.parameters()]

    def k_params(self):
        return [p for m in (self.encoder_k, self.neck_k, self.mlp_k) for p in m.parameters()]

    def extract_feat(self, im_q, im_k):
        """moco.py:517-547 with world size 1 (the shuffle draws a permutation and applies it)."""
        q_mlvl = self.encoder_q(im_q)
        (q_emb, q_mlvl), _ = self.neck_q(q_mlvl)
        q = F.normalize(self.mlp_q(q_emb), dim=1)
        with torch.no_grad():
            m = self.fixed_m if self.fixed_m is not None else O.momentum(self.state.iters, self.max_iters, self.m_base)
            for pk, new in zip(self.k_params(), O.ema_update([p.data for p in self.k_params()],
                                                             [p.data for p in self.q_params()], m)):
                pk.data = new
            idx = torch.randperm(im_k.shape[0])
            im_k, unshuf = O.batch_shuffle(im_k, idx, 0, 1)
            k_mlvl = self.encoder_k(im_k)
            (k_emb, k_mlvl), _ = self.neck_k(k_mlvl)
            k = F.normalize(self.mlp_k(k_emb), dim=1)
            k = O.batch_unshuffle(k, unshuf, 0, 1)
        return q, q_mlvl, k


def extract_feat_ranks(branch, im_qs, im_ks):
    """moco.py:517-547 for G ranks emulated in ONE process (data-parallel semantics of the reference): every rank
    encodes its own queries; the key batch of all ranks is gathered rank-major (:564-567), permuted by rank 0's
    `torch.randperm` (:160-163), each rank's key encoder sees ITS slice of the permuted batch (so batch-norm statistics
    are per rank, per shuffled subset), and the keys are gathered again and un-permuted (:174-191).
    Returns ([(q_r, q_mlvl_r)], [k_r], k_all) with k_all in rank-major order (what the enqueue receives)."""
    G, n = len(im_qs), im_qs[0].shape[0]
    outs_q = []
    for r in range(G):
        q_mlvl = branch.encoder_q(im_qs[r])
        (q_emb, q_mlvl), _ = branch.neck_q(q_mlvl)
        outs_q.append((F.normalize(branch.mlp_q(q_emb), dim=1), q_mlvl))
    with torch.no_grad():
        m = branch.fixed_m if branch.fixed_m is not None else O.momentum(branch.state.iters, branch.max_iters, branch.m_base)
        for pk, new in zip(branch.k_params(), O.ema_update([p.data for p in branch.k_params()],
                                                           [p.data for p in branch.q_params()], m)):
            pk.data = new
        idx = torch.randperm(n * G)
        x_all = torch.cat(im_ks)
        ks = []
        for r in range(G):
            sub, _ = O.batch_shuffle(x_all, idx, r, G)
            k_mlvl = branch.encoder_k(sub)
            (k_emb, _), _ = branch.neck_k(k_mlvl)
            ks.append(F.normalize(branch.mlp_k(k_emb), dim=1))
        k_all = torch.cat(ks)[torch.argsort(idx)]
    return outs_q, [k_all[r * n:(r + 1) * n] for r in range(G)], k_all


class OracleMSCL:
    def __init__(self, model, device="cpu"):
        self.rgb = OracleBranch(model.recognizer, device=device)
        self.flow = OracleBranch(model.recognizer_flow, device=device)
        self.T = model.moco_mx_head.T
        self.t = model.sup_head.labels.shape[1]
        self.mlvl_ids = model.sup_head.mlvl_ids
        self.trans_rgb = copy.deepcopy(model.sup_head.trans_rgb).to(device)   # Identity unless bkb_channels says otherwise
        self.trans_flow = copy.deepcopy(model.sup_head.trans_flow).to(device)
        self.weight_aug_flow = model.weight_aug_flow
        self.training = model.training

    def parameters(self):
        return self.rgb.q_params() + self.flow.q_params() + list(self.trans_rgb.parameters()) + list(self.trans_flow.parameters())

    def train_step(self, im_q, im_k, flow_q, flow_k):
        """Inputs already augmented: RGB (N,3,T,H,W), flow images (N,3,2T,H,W).  Returns (loss, log_vars).
        Call order is the reference's (mscl.py:227-240): each recognizer call = encoders, logits
        against the pre-enqueue snapshot, enqueue, iters."""
        from collections import OrderedDict
        tr, T = self.training, self.T
        fq, afq = (x.contiguous() for x in flow_q.chunk(2, 2))
        fk, afk = (x.contiguous() for x in flow_k.chunk(2, 2))
        q, q_mlvl, k = self.rgb.extract_feat(im_q, im_k)
        loss_img = self.rgb.state.branch(q, k, self.rgb.T, "", True, tr)
        q_f, qf_mlvl, k_f = self.flow.extract_feat(fq, fk)
        loss_flow = self.flow.state.branch(q_f, k_f, self.flow.T, "_flow", True, tr)
        q_af, qaf_mlvl, k_af = self.flow.extract_feat(afq, afk)       # EMA sees the advanced iters
        loss_aug = self.flow.state.branch(q_af, k_af, self.flow.T, "_flow", False, tr)
        loss_flow = O.merge_flow_losses(loss_flow, loss_aug, self.weight_aug_flow)
        feats = dict(q=q, k=k, q_f=q_f, k_f=k_f, q_af=q_af, k_af=k_af, q_map=q_mlvl[self.mlvl_ids[0]],
                     qf_map=qf_mlvl[self.mlvl_ids[1]], qaf_map=qaf_mlvl[self.mlvl_ids[1]])
        loss_mx, loss_sup = O.mscl_tail(feats, self.rgb.state.weight, self.flow.state.weight, T, self.t,
                                        self.weight_aug_flow, self.trans_rgb, self.trans_flow)
        losses = OrderedDict()
        for d in (loss_img, loss_flow, loss_mx, loss_sup):
            losses.update(d)
        return O.parse_losses(losses)


class OracleMoCo:
    """One MoCo / MoCoV2 recognizer trained on its own (`moco_r*.py` configs): train_step = extract_feat, logits
    against the pre-enqueue decayed snapshot, enqueue, iters (moco.py:236-296 / :473-515)."""

    def __init__(self, model):
        self.branch = OracleBranch(model)
        self.basename = model.moco_head.basename
        self.training = model.training

    def parameters(self):
        return self.branch.q_params()

    def train_step(self, im_q, im_k):
        q, _, k = self.branch.extract_feat(im_q, im_k)
        losses = self.branch.state.branch(q, k, self.branch.T, self.basename, True, self.training)
        return O.parse_losses(losses)

    def train_step_ranks(self, im_qs, im_ks):
        """The same step on G data-parallel ranks emulated in one process: every rank's logits against the SAME
        pre-enqueue snapshot, ONE enqueue of the rank-major gathered keys, iters += the global batch, log variables
        averaged over ranks (recognizers/base.py:301-306).  Returns ([loss_r], averaged log_vars)."""
        outs_q, ks, k_all = extract_feat_ranks(self.branch, im_qs, im_ks)
        st = self.branch.state
        st.weight = O.decayed_weight(st.queue, st.count)
        per_rank = []
        for (q, _), k in zip(outs_q, ks):
            logits = O.infonce_logits(q, k, st.weight, self.branch.T)
            labels = torch.zeros(logits.shape[0], dtype=torch.long)
            per_rank.append(O.parse_losses(O.head_loss(logits, labels, self.basename)))
        st.ptr = O.enqueue(st.queue, st.count, st.ptr, k_all)
        st.batch_size = k_all.shape[0]
        if self.training:
            st.iters += st.batch_size
        keys = list(per_rank[0][1].keys())
        avg = {k: sum(lv[k] for _, lv in per_rank) / len(per_rank) for k in keys}
        return [loss for loss, _ in per_rank], avg


class OracleTwoBranch:
    """MSCL (kind="mscl", recognizers/mscl.py:85-120) or MoDist (kind="modist", recognizers/modist.py:77-118) with the
    product's encoder modules on the CPU."""

    def __init__(self, model, kind):
        self.kind = kind
        self.rgb = OracleBranch(model.recognizer)
        self.flow = OracleBranch(model.recognizer_flow)
        self.training = model.training
        if kind == "mscl":
            head = model.sup_head
            self.T, self.same_kn, self.mx_basename = model.moco_mx_head.T, model.moco_mx_head.same_kn, model.moco_mx_head.basename
            self.sup = dict(t=head.labels.shape[1], T=head.T, ids=head.mlvl_ids, trans_rgb=copy.deepcopy(head.trans_rgb).cpu(),
                            trans_flow=copy.deepcopy(head.trans_flow).cpu(), with_aug=type(head).__name__ != "MoDistv2PosHead")
        else:
            self.T, self.same_kn, self.mx_basename, self.sup = model.T, model.same_kn, model.moco_head.basename, None

    def parameters(self):
        extra = []
        if self.sup is not None:
            extra = list(self.sup["trans_rgb"].parameters()) + list(self.sup["trans_flow"].parameters())
        return self.rgb.q_params() + self.flow.q_params() + extra

    def train_step(self, im_q, im_k, flow_q, flow_k):
        q, q_mlvl, k = self.rgb.extract_feat(im_q, im_k)
        # the reference runs the whole RGB call (snapshot, enqueue, iters) before the flow encoders; the enqueue does
        # not feed the flow call, so only the EMA schedule (iters) and the permutation draw order matter here
        q_f, qf_mlvl, k_f = self.flow.extract_feat(flow_q, flow_k)
        sup = None
        if self.sup is not None:
            s = self.sup
            sup = lambda: O.frame_contrast(q_mlvl[s["ids"][0]], qf_mlvl[s["ids"][1]], s["T"], s["t"], s["trans_rgb"], s["trans_flow"])
        losses = O.two_branch_objective(dict(q=q, k=k, q_f=q_f, k_f=k_f), self.rgb.state, self.flow.state, T=self.T,
                                        same_kn=self.same_kn, kind=self.kind, mx_basename=self.mx_basename, sup=sup,
                                        training=self.training)
        return O.parse_losses(losses)
