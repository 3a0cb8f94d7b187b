"""FusedClipSGD: torch.nn.utils.clip_grad_norm_ + torch.optim.SGD(momentum, weight_decay).step() as two multi-tensor
kernel launches (K10, mscl_b200/csrc/optim.cu).

The reference trains with `optimizer = dict(type='SGD', lr=0.02, momentum=0.9, weight_decay=1e-4)` and
`optimizer_config = dict(grad_clip=dict(max_norm=40, norm_type=2))`
(configs/recognition/moco/mscl_r18_cosm_lr2e-2.py:112-119; mmcv's OptimizerHook calls clip_grad_norm_ and then
optimizer.step()).  Same update rule as torch.optim.SGD (dampening 0, no Nesterov); parameters whose `.grad` is None
are skipped, as there.  A torch.optim.Optimizer subclass: `param_groups[*]['lr']` can be driven by any LR scheduler /
mmcv LrUpdaterHook, `state_dict()` holds `momentum_buffer` per parameter like SGD's.
"""
import torch

from . import _cabi
from .functional import _chk_dense, _stream

CHUNK = 16384


class FusedClipSGD(torch.optim.Optimizer):
    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, max_norm=None):
        if momentum < 0 or weight_decay < 0 or lr < 0:
            raise ValueError("lr, momentum and weight_decay must be non-negative")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self.max_norm = max_norm
        self._plans = {}
        self.last_grad_norm = None      # device tensor [2] = (||g||, clip coefficient) of the last step

    def _plan(self, gi, params):
        """Chunk tables for the parameters of one group that currently hold a gradient."""
        key = (gi, tuple(id(p) for p in params))
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        dev = params[0].device
        sizes = [p.numel() for p in params]
        blk_t, blk_s = [], []
        for i, n in enumerate(sizes):
            for s in range(0, n, CHUNK):
                blk_t.append(i)
                blk_s.append(s)
        plan = dict(n_blocks=len(blk_t), numel=sum(sizes),
                    sizes=torch.tensor(sizes, dtype=torch.int64, device=dev),
                    blk_tensor=torch.tensor(blk_t, dtype=torch.int32, device=dev),
                    blk_start=torch.tensor(blk_s, dtype=torch.int64, device=dev),
                    p_ptrs=torch.tensor([p.data_ptr() for p in params], dtype=torch.int64, device=dev),
                    partial=torch.empty(len(blk_t), device=dev), stats=torch.empty(2, device=dev),
                    p_sig=[p.data_ptr() for p in params])
        if len(self._plans) > 8:
            self._plans.clear()
        self._plans[key] = plan
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        st = _stream()
        jobs = []
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            bufs, fresh = [], []
            for p in params:
                _chk_dense(p.data, "parameter")
                if not _same_layout(p.grad, p):
                    p.grad = _restride(p.grad, p)
                _chk_dense(p.grad, "gradient")
                state = self.state[p]
                new = state.get("momentum_buffer") is None
                if new:
                    state["momentum_buffer"] = torch.empty_like(p, memory_format=torch.preserve_format)
                fresh.append(new)
                bufs.append(state["momentum_buffer"])
            if any(fresh) and not all(fresh):      # a parameter joined later: give it a zero buffer (== SGD's first step)
                for p, f in zip(params, fresh):
                    if f:
                        self.state[p]["momentum_buffer"].zero_()
            first = all(fresh)
            plan = self._plan(gi, params)
            if plan["p_sig"] != [p.data_ptr() for p in params]:
                self._plans.clear()
                plan = self._plan(gi, params)
            # gradient (and momentum-buffer) addresses change from step to step: staged through pinned memory, no host stall
            g_ptrs = torch.tensor([p.grad.data_ptr() for p in params], dtype=torch.int64).pin_memory().to(
                params[0].device, non_blocking=True)
            b_ptrs = plan.get("b_ptrs")
            if b_ptrs is None or plan.get("b_sig") != [b.data_ptr() for b in bufs]:
                plan["b_sig"] = [b.data_ptr() for b in bufs]
                b_ptrs = plan["b_ptrs"] = torch.tensor(plan["b_sig"], dtype=torch.int64, device=params[0].device)
            jobs.append((group, plan, g_ptrs, b_ptrs, first))
        if not jobs:
            return loss
        if self.max_norm is not None:
            if len(jobs) != 1:
                raise _cabi.MsclError("FusedClipSGD clips over ONE parameter group (the reference's config has one)")
            group, plan, g_ptrs, b_ptrs, first = jobs[0]
            _cabi.call("mscl_grad_norm_multi", g_ptrs.data_ptr(), plan["sizes"].data_ptr(), plan["blk_tensor"].data_ptr(),
                       plan["blk_start"].data_ptr(), plan["n_blocks"], CHUNK, float(self.max_norm), plan["partial"].data_ptr(),
                       plan["stats"].data_ptr(), st, algo_bytes=4 * plan["numel"])
            self.last_grad_norm = plan["stats"]
        for group, plan, g_ptrs, b_ptrs, first in jobs:
            _cabi.call("mscl_clip_sgd_multi", g_ptrs.data_ptr(), plan["p_ptrs"].data_ptr(), b_ptrs.data_ptr(),
                       plan["sizes"].data_ptr(), plan["blk_tensor"].data_ptr(), plan["blk_start"].data_ptr(), plan["n_blocks"],
                       CHUNK, plan["stats"].data_ptr() if self.max_norm is not None else None, float(group["weight_decay"]),
                       float(group["momentum"]), float(group["lr"]), int(first), st, algo_bytes=24 * plan["numel"])
        return loss


def _same_layout(a, b):
    """Same element order in storage (strides of size-1 dims are arbitrary and do not matter)."""
    return all(sa == sb for n, sa, sb in zip(a.shape, a.stride(), b.stride()) if n > 1)


def _restride(g, p):
    """A gradient laid out differently from its parameter (e.g. NCDHW grad of a channels_last_3d weight): same values
    in the parameter's layout, so one flat walk over both storages pairs the right elements."""
    out = torch.empty_like(p, memory_format=torch.preserve_format)
    out.copy_(g)
    return out
