"""MoCoV2 recognizer on the B200 kernels (reference: recognizers/moco.py:318-554).

Same constructor, same `train_step` / `forward_train(im_q, im_k, aux_info,
return_features, update_queue)` contract, same state_dict keys (`queue` (dim,K) fp32,
`queue_ptr` (1,) int64, `count` (K,) int64) and the same observable sequence of effects per
call: EMA -> shuffle -> key forward -> unshuffle -> logits against the
PRE-enqueue decayed queue -> enqueue -> iters.  What changed is how each step runs:

  _momentum_update_key_encoder   one multi-tensor kernel (K4) instead of ~3 launches per tensor
  _batch_shuffle_ddp             permutation-driven all-to-all (K6) instead of gather-everything
  logits / CE / top-k            one fused tcgen05 pass over the queue (K1); no (N,1+K) matrix,
                                 no decayed snapshot, no host argsort
  _dequeue_and_enqueue           one block copy into a key-major ring buffer with implicit ages (K5)
"""
from math import cos, pi

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as fx
from ..registry import RECOGNIZERS, build_ssl_aug
from .base_moco import BaseMoCoRecognizer
from . import shuffle as shf


def _dist_on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


@torch.no_grad()
def concat_all_gather(tensor):
    """Rank-major all_gather without gradient (moco.py:558-568)."""
    if not _dist_on():
        return tensor
    out = torch.empty((dist.get_world_size() * tensor.shape[0],) + tuple(tensor.shape[1:]),
                      dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out, tensor.contiguous())
    return out


@RECOGNIZERS.register_module()
class MoCoV2(BaseMoCoRecognizer):
    def __init__(self, backbone, neck, moco_head, im_key="imgs", dim_in=512, dim=128, K=65536, m_base=0.994,
                 t_decay=0.99999, max_iters=1, T=0.07, mlp=False, aux_info=[],
                 aug=dict(dtype="MoCoAugmentV3", moco_aug=(112, 112), t=8), train_cfg=None, test_cfg=None):
        super().__init__(train_cfg=train_cfg, test_cfg=test_cfg)
        self.K = K
        self.m_base = m_base
        self.m = m_base
        self.iters = 0
        self.max_iters = max_iters
        self.batch_size = 0
        self.T = T
        self.dim = dim
        self.im_key = im_key
        self.t_decay = t_decay      # stored but unused, like the reference (the decay base is 0.99999)
        self.aux_info = aux_info
        if "num_classes" in backbone:
            assert backbone["num_classes"] == dim
        self._build_backbone(backbone, "encoder_q")
        self._build_backbone(dict(backbone), "encoder_k")
        self._build_neck(neck, "neck_q")
        self._build_neck(neck, "neck_k")
        self._build_cls_head(moco_head, "moco_head")
        self.init_weights()
        if mlp:
            self.mlp_q = nn.Sequential(nn.Linear(dim_in, dim_in), nn.ReLU(), nn.Linear(dim_in, dim))
            self.mlp_k = nn.Sequential(nn.Linear(dim_in, dim_in), nn.ReLU(), nn.Linear(dim_in, dim))
        else:
            self.mlp_q = nn.Linear(dim_in, dim)
            self.mlp_k = nn.Linear(dim_in, dim)
        for mq, mk in self._qk_modules():
            for pq, pk in zip(mq.parameters(), mk.parameters()):
                pk.data.copy_(pq.data)
                pk.requires_grad = False
        # the queue starts as unit-norm Gaussian columns with age 0 (moco.py:390-396); it is staged
        # on the host until the module is first used on a device
        queue0 = F.normalize(torch.randn(dim, K), dim=0)
        self._staged = dict(queue=queue0, count=torch.zeros(K, dtype=torch.long), ptr=0)
        self._nq = None
        self._ema = None
        self._weight = None
        self._cpu_group = None
        cfg = train_cfg or {}
        self.shard_queue = bool(cfg.get("shard_queue", False))
        self.keep_weight_snapshot = bool(cfg.get("keep_weight_snapshot", False))
        self.aug_gpu = build_ssl_aug(aug)
        self._register_state_dict_hook(MoCoV2._export_queue_hook)
        self._register_load_state_dict_pre_hook(self._import_queue_hook)

    def _qk_modules(self):
        return ((self.encoder_q, self.encoder_k), (self.neck_q, self.neck_k), (self.mlp_q, self.mlp_k))

    # ------------------------------------------------------------------ queue state
    def negative_queue(self, device=None):
        """The device ring buffer, created (or moved) on first use."""
        if device is None:
            device = next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise fx._cabi.MsclError("MoCoV2 needs its parameters on a CUDA device (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._nq is not None and self._nq.device == device:
            return self._nq
        if self._nq is not None:       # module was moved: carry the state over
            self._staged = self._gathered_state()
        rank = dist.get_rank() if _dist_on() else 0
        world = dist.get_world_size() if _dist_on() else 1
        nq = fx.NegativeQueue(self.K, self.dim, device, rank, world, shard=self.shard_queue and world > 1)
        if world > 1:
            # The reference registers queue / count / queue_ptr as buffers, so DDP's initial module-state sync makes every
            # rank start from rank 0's random queue (moco.py:390-396); here they are plain state staged on the host and
            # drawn per rank, so rank 0's copy is broadcast on first use (a collective: every rank's first forward).
            st = self._staged
            q0 = st["queue"].to(device, torch.float32).contiguous()
            c0 = st["count"].to(device, torch.int64).contiguous()
            p0 = torch.tensor([int(st["ptr"])], dtype=torch.int64, device=device)
            for t in (q0, c0, p0):
                dist.broadcast(t, src=0)
            self._staged = dict(queue=q0, count=c0, ptr=int(p0.item()))
        with torch.cuda.device(device):
            nq.load(self._staged["queue"], self._staged["count"], self._staged["ptr"])
        self._nq, self._staged = nq, None
        return nq

    def _gathered_state(self):
        """Reference-layout state (queue (C,K), count (K,), ptr).  NOT a collective, sharded queue or not: every rank
        keeps the fp32 master of the whole queue (functional.NegativeQueue), so `state_dict()` may be called from one
        rank alone -- mmcv's checkpoint hook runs under `@master_only`."""
        if self._nq is None:
            return self._staged
        q, c = self._nq.export_full()
        return dict(queue=q, count=c, ptr=self._nq.ptr)

    @property
    def queue(self):
        return self._gathered_state()["queue"]

    @property
    def count(self):
        return self._gathered_state()["count"]

    @property
    def queue_ptr(self):
        st = self._gathered_state()
        return torch.tensor([st["ptr"]], dtype=torch.long, device=st["queue"].device)

    @staticmethod
    def _export_queue_hook(module, state_dict, prefix, local_metadata):
        st = module._gathered_state()
        state_dict[prefix + "queue"] = st["queue"].detach().clone()
        state_dict[prefix + "queue_ptr"] = torch.tensor([st["ptr"]], dtype=torch.long, device=st["queue"].device)
        state_dict[prefix + "count"] = st["count"].detach().clone()
        return state_dict

    def _import_queue_hook(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        keys = [prefix + k for k in ("queue", "queue_ptr", "count")]
        if not all(k in state_dict for k in keys):
            if strict:
                missing_keys.extend(k for k in keys if k not in state_dict)
            for k in keys:
                state_dict.pop(k, None)
            return
        queue, ptr, count = (state_dict.pop(k) for k in keys)
        if tuple(queue.shape) != (self.dim, self.K) or tuple(count.shape) != (self.K,):
            error_msgs.append(f"size mismatch for {prefix}queue/count: got {tuple(queue.shape)} / {tuple(count.shape)}")
            return
        staged = dict(queue=queue.detach().float().cpu(), count=count.detach().long().cpu(), ptr=int(ptr.view(-1)[0]))
        if self._nq is not None:
            dev = self._nq.device
            self._nq = None
            self._staged = staged
            self.negative_queue(dev)
        else:
            self._staged = staged

    @property
    def weight(self):
        """Decayed queue snapshot (C,K) of the last forward_train, as the reference keeps in
        `self._weight` (moco.py:488,549-551).  Materialised only on request."""
        if self._weight is None:
            raise RuntimeError("no snapshot kept: call forward_train(..., return_features=True) or set "
                               "train_cfg=dict(keep_weight_snapshot=True)")
        return self._weight

    # ------------------------------------------------------------------ K4
    @torch.no_grad()
    def _next_momentum(self):
        """Cosine-annealed momentum of this update (moco.py:413-415)."""
        factor = min(self.iters / self.max_iters, 1)
        return 1 * (1 - 0.5 * (1 - self.m_base) * (cos(pi * factor) + 1))

    def _momentum_update_key_encoder(self):
        self.m = self._next_momentum()
        if self._ema is None:
            pk, pq = [], []
            for mq, mk in self._qk_modules():
                pq += list(mq.parameters())
                pk += list(mk.parameters())
            self._ema = fx.EmaTable(pk, pq)
        self._ema.update(self.m)

    # ------------------------------------------------------------------ K5
    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys, save=False):
        """moco.py:423-440.  save=True: returns what the enqueue overwrote (functional.NegativeQueue.enqueue)."""
        keys = concat_all_gather(keys.contiguous())
        self.batch_size = keys.shape[0]
        return self.negative_queue(keys.device).enqueue(keys.contiguous(), save=save)

    # ------------------------------------------------------------------ K6
    def _shuffle_group(self):
        if self._cpu_group is None and _dist_on():
            self._cpu_group = dist.new_group(backend="gloo")
        return self._cpu_group

    @torch.no_grad()
    def _batch_shuffle_ddp(self, x):
        """Returns (rows this rank feeds its key encoder, idx_unshuffle) -- moco.py:146-172."""
        world = dist.get_world_size() if _dist_on() else 1
        n = x.shape[0]
        idx_shuffle = shf.draw_permutation(n * world, x.device, self._shuffle_group())
        idx_unshuffle = torch.argsort(idx_shuffle)
        if world == 1:
            # one rank: the key encoder sees the same set of clips whatever their order, and the
            # unshuffle restores it; batch-norm statistics are permutation invariant -> no copy
            return x, idx_unshuffle
        plan = shf.ShufflePlan(idx_shuffle, n, dist.get_rank(), world)
        return shf.exchange(x.contiguous(), plan, fx.gather_rows), idx_unshuffle

    @torch.no_grad()
    def _batch_shuffle_begin(self, x):
        """The same shuffle with the all-to-all left in flight: returns (pending, idx_unshuffle); `pending.finish()` gives
        the rows this rank feeds its key encoder.  Only used with more than one rank."""
        world = dist.get_world_size()
        n = x.shape[0]
        idx_shuffle = shf.draw_permutation(n * world, x.device, self._shuffle_group())
        plan = shf.ShufflePlan(idx_shuffle, n, dist.get_rank(), world)
        return shf.exchange_begin(x.contiguous(), plan, fx.gather_rows), torch.argsort(idx_shuffle)

    @torch.no_grad()
    def _batch_unshuffle_ddp(self, x, idx_unshuffle):
        """moco.py:174-191."""
        world = dist.get_world_size() if _dist_on() else 1
        if world == 1:
            return x
        plan = shf.ShufflePlan(idx_unshuffle, x.shape[0], dist.get_rank(), world)
        return shf.exchange(x.contiguous(), plan, fx.gather_rows)

    # ------------------------------------------------------------------ encoders
    def q_path(self, im_q):
        """Query side: encoder -> neck -> projection -> L2 norm (moco.py:520-529).  Returns (q, q_mlvl, sup_loss)."""
        q_mlvl = self.encoder_q(im_q)
        (q_emb, q_mlvl), sup_loss = self.neck_q(q_mlvl)
        return F.normalize(self.mlp_q(q_emb), dim=1), q_mlvl, sup_loss

    def k_path(self, im_k, pyramid=True):
        """Key side (moco.py:538-541), no gradient.  Returns (k, k_mlvl)."""
        k_mlvl = self.encoder_k(im_k)
        (k_emb, k_mlvl), _ = self.neck_k(k_mlvl, pyramid=pyramid)
        return F.normalize(self.mlp_k(k_emb), dim=1), k_mlvl

    def extract_feat(self, im_q, im_k, unshuffle_mlvl=True, site=0):
        """Returns q, q_mlvl, k, k_mlvl, sup_loss (moco.py:517-547).  `site` names the call site within a step (the
        flow recognizer is called twice): CUDA-graphed encoder paths (mscl_b200/graphed.py) keep one set of static
        activations per site."""
        graphed = getattr(self, "_graphed_paths", None)
        pending = None
        if _dist_on() and dist.get_world_size() > 1:
            # More than one rank: the key encoder's momentum update and the shuffle exchange are started BEFORE the query
            # forward instead of after it (moco.py:531-536), so the all-to-all travels while the query encoder runs.  Same
            # results: the update reads the query parameters, which the forward does not change, and the permutation is
            # still the next draw of the CPU generator (nothing in the query forward touches it).
            with torch.no_grad():
                self._momentum_update_key_encoder()
                pending, idx_unshuffle = self._batch_shuffle_begin(im_k)
        if graphed is not None and self.training and torch.is_grad_enabled():
            q, q_mlvl, sup_loss = graphed.q(site, im_q)
        else:
            q, q_mlvl, sup_loss = self.q_path(im_q)
        with torch.no_grad():
            if pending is not None:
                im_k = pending.finish()
            else:
                self._momentum_update_key_encoder()
                im_k, idx_unshuffle = self._batch_shuffle_ddp(im_k)
            if graphed is not None and self.training and not unshuffle_mlvl:
                k, k_mlvl = graphed.k(site, im_k)
            else:
                k, k_mlvl = self.k_path(im_k, pyramid=unshuffle_mlvl)
            k = self._batch_unshuffle_ddp(k, idx_unshuffle)
            if unshuffle_mlvl:     # never consumed by MSCLWithAug: skipped there
                k_mlvl = [self._batch_unshuffle_ddp(lvl, idx_unshuffle) for lvl in k_mlvl]
        return q, q_mlvl, k, k_mlvl, sup_loss

    # ------------------------------------------------------------------ the objective
    def _group(self):
        nq = self._nq
        return dist.group.WORLD if (nq is not None and nq.world > 1) else None

    def _stack_terms(self, terms):
        """(q, k_pos[, dup_slot]) row sets -> stacked q, k_pos, dup_slot (or None) and the rows per term."""
        n = terms[0][0].shape[0]
        q = torch.cat([t[0] for t in terms], dim=0).contiguous()
        kp = torch.cat([t[1].detach() for t in terms], dim=0).contiguous()
        dups = [t[2] if len(t) > 2 else None for t in terms]
        dup = None
        if any(d is not None for d in dups):
            none = torch.full((n,), -1, dtype=torch.int32, device=q.device)
            dup = torch.cat([none if d is None else d for d in dups]).contiguous()
        return q, kp, dup, n

    def contrast(self, terms, T=None):
        """Fused InfoNCE of several (q, k_pos[, dup_slot]) row sets against the CURRENT queue state in
        one pass.  dup_slot: int32 (n,) global slots holding copies of the term's own positive keys
        (a term whose keys were enqueued before the pass), or None.
        Returns a (len(terms), 4) tensor of [loss, top1, top5, 0] rows."""
        q, kp, dup, n = self._stack_terms(terms)
        nq = self.negative_queue(q.device)
        out, _ = fx.infonce(q, kp, nq, n, self.T if T is None else T, group=self._group(), dup_slot=dup)
        return out

    @staticmethod
    def can_launch_together(recs, device):
        """Can `contrast_many` put passes over these recognizers' queues into ONE launch?  (CUDA, distinct unsharded
        queues, at most four.)"""
        if torch.device(device).type != "cuda" or not 1 <= len(recs) <= 4:
            return False
        queues = [rec.negative_queue(device) for rec in recs]
        return all(nq.world == 1 for nq in queues) and len({id(nq) for nq in queues}) == len(queues)

    @staticmethod
    def contrast_many(calls):
        """Several `contrast` calls on DIFFERENT recognizers (queues) that do not depend on each other, as ONE launch
        when none of the queues is sharded (functional.infonce_multi: the jobs share the launch's fixed costs);
        otherwise one after the other.  calls: list of (recognizer, terms, T) or (recognizer, terms, T, split) with
        split = (overwritten, n_pre): the first n_pre terms read the recognizer's queue as it was BEFORE its last enqueue
        (`overwritten` = what that `_dequeue_and_enqueue(..., save=True)` returned), the others as it is -- one call at
        most, and only where `can_launch_together` holds.  Returns the list of their results."""
        calls = [tuple(c) + (None,) * (4 - len(c)) for c in calls]
        has_split = any(sp is not None for _, _, _, sp in calls)
        together = all(terms[0][0].is_cuda for _, terms, _, _ in calls) and \
            MoCoV2.can_launch_together([rec for rec, _, _, _ in calls], calls[0][1][0][0].device)
        if has_split and not together:
            raise fx._cabi.MsclError("an epoch-split pass needs unsharded CUDA queues (check can_launch_together first)")
        if not together or (len(calls) < 2 and not has_split):
            return [rec.contrast(terms, T) for rec, terms, T, _ in calls]      # (contrast refuses host tensors itself)
        jobs = []
        for rec, terms, T, sp in calls:
            q, kp, dup, n = rec._stack_terms(terms)
            job = dict(q=q, kpos=kp, nq=rec.negative_queue(q.device), rows_per_group=n, T=rec.T if T is None else T, dup_slot=dup)
            if sp is not None:
                job.update(overwritten=sp[0], row_split=int(sp[1]) * n)
            jobs.append(job)
        return [out for out, _ in fx.infonce_multi(jobs)]

    def enqueue_slots(self, n_local, device):
        """Global queue slots the NEXT enqueue writes this rank's n_local keys to (rank-major gather
        order, moco.py:434,564-567), as int32 (n_local,)."""
        nq = self.negative_queue(device)
        rank = dist.get_rank() if _dist_on() else 0
        return torch.arange(n_local, dtype=torch.int32, device=device) + int(nq.ptr + rank * n_local)

    def note_branch(self, n_local, update_queue=True):
        """Host-side bookkeeping of one forward_train call (moco.py:429,504-505): the gathered batch
        size is recorded by the enqueue, then `iters` advances by it while training."""
        if update_queue:
            self.batch_size = n_local * (dist.get_world_size() if _dist_on() else 1)
        if self.training:
            self.iters += self.batch_size

    def after_branch(self, k, update_queue=True):
        """Enqueue + iteration bookkeeping of one forward_train call (moco.py:500-505)."""
        if update_queue:
            self._dequeue_and_enqueue(k)
        self.note_branch(k.shape[0], False)

    def train_step(self, data_batch, optimizer, **kwargs):
        im_q = data_batch[self.im_key][0]
        im_k = data_batch[self.im_key][1]
        aux_info = {}
        for item in self.aux_info:
            assert item in data_batch
            aux_info[item] = data_batch[item]
        losses = self(im_q, im_k, aux_info, return_loss=True)
        loss, log_vars = self._parse_losses(losses)
        return dict(num_samples=im_q.shape[0], loss=loss, log_vars=log_vars)

    def forward(self, im_q, im_k, aux_info, return_loss=True, **kwargs):
        if kwargs.get("gradcam", False):
            del kwargs["gradcam"]
            return self.forward_gradcam(im_q, im_k, aux_info, **kwargs)
        if return_loss:
            return self.forward_train(im_q, im_k, aux_info, **kwargs)
        raise NotImplementedError("MoCo doesnt support test mode")

    def forward_train(self, im_q, im_k, aux_info, return_features=False, update_queue=True):
        if return_features:
            aux_info = aux_info.copy()
        else:
            im_q, im_k, aux_info = self.aug_gpu(im_q, im_k, aux_info)
        q, q_mlvl, k, k_mlvl, sup_loss = self.extract_feat(im_q, im_k)
        if not self.moco_head.can_fuse():
            raise NotImplementedError("the fused path needs loss_cls=CrossEntropyLoss_torch without class weights")
        out = self.contrast([(q, k)])                        # snapshot semantics: before the enqueue
        if return_features or self.keep_weight_snapshot:    # an outer recognizer may read `.weight`
            self._weight = self._gathered_weight()
        self.after_branch(k, update_queue)
        losses = self.moco_head.loss_fused(out[0])
        losses.update(sup_loss)
        if return_features:
            return losses, dict(q=q, q_mlvl=q_mlvl, k=k, k_mlvl=k_mlvl, q_neg=None)
        return losses

    def _gathered_weight(self):
        w = self._nq.weight()
        if self._nq.world > 1:
            ws = [torch.empty_like(w) for _ in range(self._nq.world)]
            dist.all_gather(ws, w)
            w = torch.cat(ws, dim=1)
        return w

    def visualize(self, data_batch):
        pass


@RECOGNIZERS.register_module()
class MoCo(MoCoV2):
    """The reference's first MoCo recognizer (recognizers/moco.py:30-316): everything MoCoV2 does -- decay-by-age
    negatives included (:270-273) -- with a CONSTANT key-encoder momentum `m` (:114-124) instead of the cosine
    schedule.  Used by the four `moco_r*.py` configs.  (In the reference MoCoV2 derives from MoCo; here the
    machinery lives in MoCoV2 and this class only pins the momentum.)"""

    def __init__(self, backbone, neck, moco_head, im_key="imgs", dim_in=512, dim=128, K=65536, m=0.999, T=0.07,
                 mlp=False, aux_info=[], aug=dict(dtype="MoCoAugmentV3", moco_aug=(112, 112), t=8),
                 train_cfg=None, test_cfg=None):
        super().__init__(backbone, neck, moco_head, im_key=im_key, dim_in=dim_in, dim=dim, K=K, m_base=m, max_iters=1,
                         T=T, mlp=mlp, aux_info=aux_info, aug=aug, train_cfg=train_cfg, test_cfg=test_cfg)
        self.m = m

    def _next_momentum(self):
        return self.m
