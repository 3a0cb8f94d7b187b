"""Shuffle-BN data movement (MoCo._batch_shuffle_ddp / _batch_unshuffle_ddp,
mmaction/models/recognizers/moco.py:146-191).

The reference all-gathers every rank's clips (G x 38.5 MB), draws
`torch.randperm(B_all)` on the CPU default generator on every rank, broadcasts rank 0's
draw and keeps N of the G*N gathered rows.  Here the permutation source is identical
(same generator, same draw order, rank 0 wins) but only the rows a rank actually needs
travel: an all-to-all whose send/receive plan is computed on the host from the
permutation.  The result is the same tensor the reference produces.

`ShufflePlan` is pure index arithmetic (CPU tensors) and is what the gloo tests exercise.
"""
import torch
import torch.distributed as dist


class ShufflePlan:
    """Routing of `x_gather[idx_this]` for one rank without materialising x_gather.

    want[r]  = idx.view(G, -1)[r]            global rows rank r must end up with, in order
    global row g lives on rank g // n at local position g % n (rank-major gather order,
    moco.py:564-567).
    """

    def __init__(self, idx, n_local, rank, world):
        idx = idx.cpu().view(world, -1)
        assert idx.shape[1] == n_local, "the permutation must cover world * n_local rows"
        self.n, self.rank, self.world = n_local, rank, world
        send_rows, self.in_splits = [], []
        for dst in range(world):               # rows of MINE that dst wants, in dst's order
            want = idx[dst]
            mine = want[(want // n_local) == rank] % n_local
            send_rows.append(mine)
            self.in_splits.append(int(mine.numel()))
        self.send_idx = torch.cat(send_rows) if send_rows else torch.empty(0, dtype=torch.long)
        want = idx[rank]
        src = want // n_local
        self.out_splits = [int((src == s).sum()) for s in range(world)]
        # received buffer is ordered by source rank, then by the order the source sent (= my order)
        order = torch.argsort(src, stable=True)              # positions of `want` grouped by source
        self.recv_pos = torch.empty_like(order)
        self.recv_pos[order] = torch.arange(order.numel())   # want position -> index in recv buffer


def draw_permutation(b_all, device, group=None):
    """torch.randperm on the CPU default generator on every rank, rank 0's draw wins
    (moco.py:160-163).  Returns the CPU permutation."""
    idx = torch.randperm(b_all)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if group is not None:            # CPU-side group (gloo): no device round trip
            dist.broadcast(idx, src=0, group=group)
        else:
            dev_idx = idx.to(device)
            dist.broadcast(dev_idx, src=0)
            idx = dev_idx.cpu()
    return idx


def _upload(idx, dev):
    """Host index tensor -> device without stalling the host: a pageable source makes the copy synchronous,
    which would stop the CPU from running ahead of the GPU six times per step."""
    if dev.type != "cuda":
        return idx.to(dev)
    # ... and not through the copy engine either, where the few hundred bytes would wait behind the loader's next batch
    from .. import functional as fx
    return fx.host_stage(dev).upload([idx.contiguous().numpy()], dev)[0].view(idx.shape)


class PendingExchange:
    """An exchange whose all-to-all is in flight (async): `finish()` waits for it and orders the received rows."""

    def __init__(self, work, recv, send, plan, gather_rows):
        self.work, self.recv, self.send, self.plan, self.gather_rows = work, recv, send, plan, gather_rows

    def finish(self):
        self.work.wait()                     # the current stream now waits for the collective; the host does not
        return self.gather_rows(self.recv, _upload(self.plan.recv_pos, self.recv.device))


def exchange_begin(x, plan, gather_rows):
    """First half of `exchange`: gather the rows to send and START the all-to-all (async_op).  Whatever the caller launches
    before `finish()` overlaps with the transfer -- MoCoV2.extract_feat runs the query encoder there, so the 38 MB of clips
    a rank trades per shuffle no longer sit exposed in front of the key encoder."""
    send = gather_rows(x, _upload(plan.send_idx, x.device))
    recv = torch.empty_like(x)
    work = dist.all_to_all_single(recv, send, output_split_sizes=plan.out_splits, input_split_sizes=plan.in_splits, async_op=True)
    return PendingExchange(work, recv, send, plan, gather_rows)


def exchange(x, plan, gather_rows):
    """Run a ShufflePlan on device tensor x (n_local, ...) with torch.distributed all_to_all.
    `gather_rows(x, idx)` is the row-gather kernel (functional.gather_rows)."""
    dev = x.device
    send = gather_rows(x, _upload(plan.send_idx, dev))
    recv = torch.empty_like(x)
    dist.all_to_all_single(recv, send, output_split_sizes=plan.out_splits, input_split_sizes=plan.in_splits)
    return gather_rows(recv, _upload(plan.recv_pos, dev))
