"""CUDA-graphed encoder paths.

After the kernels of this repo and the NDHWC layout, the r18 step is ~32 ms of device work issued as ~4500 launches:
the host needs about as long to issue them as the GPU to run them.  The encoders (cuDNN convolutions, batch norms,
element-wise glue: ~85 % of the launches) have static shapes, so each call site of

    query side   encoder_q -> neck_q -> mlp_q -> normalise      (forward AND backward, torch.cuda.make_graphed_callables)
    key side     encoder_k -> mlp_k -> normalise                (forward only, one torch.cuda.CUDAGraph each)

is captured once and replayed.  Everything with per-step host state (EMA momentum, queue pointer, shuffle permutation,
the fused objective) stays eager.  The flow recognizer runs twice per step (base flow, FRA flow): every call site gets
its own graph and static activations over the SAME parameters.

    paths = graphed.enable(model, im_q, flow_q)     # model: MSCLWithAug on a CUDA device, in train mode
    ...train as usual...
    paths.after_backward()                          # re-None the gradients PyTorch eager would not have produced

Semantics kept: parameters a loss does not reach get `grad = None` in eager PyTorch, so SGD skips them (no weight
decay).  A graphed backward returns zeros for them instead; `after_backward()` restores None for the set recorded from
one eager step.
"""
import torch
import torch.nn as nn


class _QSite(nn.Module):
    """One call site of a recognizer's query path; shares the recognizer's modules (and parameters)."""

    def __init__(self, rec):
        super().__init__()
        self.encoder, self.neck, self.mlp = rec.encoder_q, rec.neck_q, rec.mlp_q
        self._rec = [rec]          # not a submodule: no parameter duplication in .parameters()

    def forward(self, x):
        q, q_mlvl, sup_loss = self._rec[0].q_path(x)
        if sup_loss:
            raise NotImplementedError("graphed query path: the neck returned an auxiliary loss")
        return (q,) + tuple(q_mlvl)


def _static_like(rec, sample):
    """A static input buffer for `sample` in the layout the encoder's first convolution wants: when the weights are
    channels_last_3d the per-step `static.copy_(clip)` then does the NCDHW -> NDHWC transposition in the one copy that
    is needed anyway, instead of a second conversion kernel inside every replayed graph."""
    w = next(rec.encoder_q.parameters())
    cl = w.dim() == 5 and w.is_contiguous(memory_format=torch.channels_last_3d) and not w.is_contiguous()
    fmt = torch.channels_last_3d if (cl and sample.dim() == 5) else torch.contiguous_format
    return sample.detach().clone(memory_format=fmt)


class _KSite:
    """Forward-only graph of a recognizer's key path (no pyramid: MSCLWithAug never reads k_mlvl)."""

    def __init__(self, rec, sample):
        self.rec = rec
        self.static_in = _static_like(rec, sample)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                rec.k_path(self.static_in, pyramid=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_k, _ = rec.k_path(self.static_in, pyramid=False)

    def __call__(self, x):
        self.static_in.copy_(x)
        self.graph.replay()
        return self.static_k.clone(), []


class GraphedPaths:
    def __init__(self, rec, samples_q, samples_k):
        self.rec = rec
        sites = [_QSite(rec) for _ in samples_q]
        for s in sites:
            s.train(rec.training)
        graphed = torch.cuda.make_graphed_callables(tuple(sites), tuple((_static_like(rec, x),) for x in samples_q),
                                                    allow_unused_input=True)   # e.g. pyramid convs of a neck whose levels feed nothing
        self.q_sites = list(graphed) if isinstance(graphed, (tuple, list)) else [graphed]
        self.k_sites = [_KSite(rec, x) for x in samples_k]

    def q(self, site, x):
        out = self.q_sites[site](x)
        return out[0], list(out[1:]), {}

    def k(self, site, x):
        return self.k_sites[site](x)


class Enabled:
    def __init__(self, model, unused):
        self.model, self.unused = model, unused

    def after_backward(self):
        for p in self.unused:
            p.grad = None


def enable(model, im_q, flow_q, eager_step=None):
    """Graph the encoder paths of an MSCLWithAug model.  im_q: an RGB clip batch as the encoders see it (after the
    augmentation), flow_q: a 3-channel flow-image batch of ONE half (base or FRA) -- shapes, dtypes and memory
    formats must be the ones of the training loop.  eager_step(): runs one ordinary forward+backward (grads left in
    place) so the set of parameters the loss does not reach can be recorded; None skips that bookkeeping."""
    if not (im_q.is_cuda and flow_q.is_cuda):
        raise RuntimeError("graphed encoder paths need CUDA tensors")
    unused = []
    if eager_step is not None:
        for p in model.parameters():
            p.grad = None
        eager_step()
        unused = [p for p in model.parameters() if p.requires_grad and p.grad is None]
        for p in model.parameters():
            p.grad = None
    rec, recf = model.recognizer, model.recognizer_flow
    rec._graphed_paths = GraphedPaths(rec, [im_q], [im_q])
    recf._graphed_paths = GraphedPaths(recf, [flow_q, flow_q], [flow_q, flow_q])
    return Enabled(model, unused)


def disable(model):
    for rec in (model.recognizer, model.recognizer_flow):
        rec._graphed_paths = None
