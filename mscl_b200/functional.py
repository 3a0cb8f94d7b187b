"""PyTorch-facing ops of the MSCL contrastive hot path.

Each op takes CUDA fp32 tensors, hands raw pointers + the current stream to the C-ABI
(include/mscl_b200.h) and returns tensors; the differentiable ones are
torch.autograd.Functions.  PyTorch is plumbing here (memory, streams, autograd graph,
torch.distributed); all arithmetic on the path runs in the sm_100a kernels.  There is
no CPU path: a non-CUDA tensor or a missing library raises.
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi

DIM = 128          # MSCL_DIM
PACK_LD = 132      # MSCL_PACK_LD


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype=torch.float32, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _cabi.MsclError(f"{name} must be a CUDA tensor (the MSCL hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise _cabi.MsclError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _cabi.MsclError(f"{name} must be contiguous")
    _cabi.require_device(t.device.index if t.device.index is not None else torch.cuda.current_device())
    return t


def _chk_dense(t, name):
    """Like _chk, but any dense layout (row-major, channels_last, channels_last_3d) passes: the
    elementwise EMA walks the storage, so only density and equal strides of the pair matter."""
    if t.is_contiguous():
        return _chk(t, name=name)
    dense = (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)) or \
            (t.dim() == 5 and t.is_contiguous(memory_format=torch.channels_last_3d))
    if not (t.is_cuda and t.dtype == torch.float32 and dense):
        raise _cabi.MsclError(f"{name} must be a dense CUDA fp32 tensor")
    return t


def _is_cl3d(t):
    """Dense channels_last_3d (NDHWC in memory) and not simultaneously row-major (C == 1 or T*H*W == 1 are both)."""
    return t.dim() == 5 and t.is_contiguous(memory_format=torch.channels_last_3d) and not t.is_contiguous()


_SM_COUNT = {}


def sm_count(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


class HostStage:
    """A ring of small pinned host buffers whose contents reach the device through `mscl_fetch_host` (a kernel reading the
    mapped buffer), not through the copy engine: per-step parameters and index tables produced on the host arrive in stream
    order without queueing behind a data loader's bulk H2D copies.  A slot is reused only after the fetch that read it has
    run."""

    def __init__(self, slots=64):      # several steps of uploads: reusing a slot must never wait for the step in flight
        self.slots = [None] * slots
        self.events = [None] * slots
        self.i = 0

    def upload(self, arrays, device):
        """arrays: list of NumPy arrays (any fixed-size dtype) -> list of flat device tensors of the same dtypes (views of
        one device buffer, each starting on a 16-byte boundary), in ONE launch."""
        arrays = [np.ascontiguousarray(a) for a in arrays]
        offs, total = [], 0
        for a in arrays:
            offs.append(total)
            total += (a.nbytes + 15) // 16 * 16
        total = max(total, 16)
        s = self.i
        self.i = (s + 1) % len(self.slots)
        if self.events[s] is not None:
            self.events[s].synchronize()            # the fetch that last read this slot has run (long ago, normally)
        if self.slots[s] is None or self.slots[s].numel() < total:
            self.slots[s] = torch.zeros(max(total, 4096), dtype=torch.uint8, pin_memory=True)
        host = self.slots[s].numpy()
        for a, o in zip(arrays, offs):
            host[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
        dev = torch.empty(total, dtype=torch.uint8, device=device)
        _cabi.call("mscl_fetch_host", dev.data_ptr(), self.slots[s].data_ptr(), total, _stream())
        if self.events[s] is None:
            self.events[s] = torch.cuda.Event()
        self.events[s].record()
        return [dev[o:o + a.nbytes].view(torch.from_numpy(a[:0]).dtype) for a, o in zip(arrays, offs)]


_STAGES = {}


def host_stage(device):
    """The calling thread's HostStage of `device`."""
    import threading
    key = (threading.get_ident(), str(device))
    st = _STAGES.get(key)
    if st is None:
        st = _STAGES[key] = HostStage()
    return st


# ----------------------------------------------------------------------------------------
# K5: the negative queue
# ----------------------------------------------------------------------------------------
class NegativeQueue:
    """Device-resident ring buffer of negatives with implicit ages.

    Replaces the reference's `queue` (C,K) fp32 / `queue_ptr` / `count` buffers
    (mmaction/models/recognizers/moco.py:390-396) on the hot path.  Layout: key-major
    fp32 [K_local, C]; ages as int32 `birth` with count = n_enq - birth; `qstate`
    int64[4] = {ptr, n_enq, counter, 0} lives on the device so no call ever syncs.
    With world_size G > 1 and shard=True each rank owns K/G consecutive slots: that is what the
    tensor-core pass streams.  Every rank ALSO keeps the bit-exact fp32 master of the WHOLE queue
    (`full_queue`, `full_birth`; 32 MB at K = 65536 -- the gathered keys every rank already holds
    are written into it by a second enqueue launch), so that state_dict() / `.queue` / `.count`
    never need a collective: mmcv saves checkpoints from rank 0 only (`@master_only`), where a
    gather of the shards would hang.
    """

    def __init__(self, K, C=DIM, device="cuda", rank=0, world=1, shard=False):
        if C != DIM:
            raise _cabi.MsclError(f"feature dim must be {DIM}")
        self.K, self.C = int(K), int(C)
        self.world = world if shard else 1
        self.rank = rank if shard else 0
        if self.K % self.world:
            raise _cabi.MsclError(f"K={K} is not divisible by the number of shards {self.world}")
        self.K_local = self.K // self.world
        self.shard_begin = self.rank * self.K_local
        self.device = torch.device(device)
        self.queue = torch.zeros(self.K_local, C, device=self.device)        # bit-exact fp32 master
        self.queue_tf32 = torch.zeros(self.K_local, C, device=self.device)   # RN-rounded operand copy
        self.birth = torch.zeros(self.K_local, dtype=torch.int32, device=self.device)
        self.qstate = torch.zeros(4, dtype=torch.int64, device=self.device)
        if self.world > 1:
            self.full_queue = torch.zeros(self.K, C, device=self.device)
            self.full_birth = torch.zeros(self.K, dtype=torch.int32, device=self.device)
            self.full_qstate = torch.zeros(4, dtype=torch.int64, device=self.device)
        self.ptr = 0          # host mirrors (deterministic, never read back from the device)
        self.n_enq = 0
        self.max_key_norm = 1.0
        self._fresh = True    # written since the last pass over it (see _prefetch_flag)

    # -- reference-layout import / export (state_dict compatibility) --
    def load(self, queue_ck, count, ptr):
        """queue_ck: (C, K) fp32 full queue, count: (K,) int64, ptr: int (moco.py:390-396)."""
        queue_ck = queue_ck.to(self.device, torch.float32)
        count = count.to(self.device, torch.int64)
        sl = slice(self.shard_begin, self.shard_begin + self.K_local)
        q_loc = queue_ck[:, sl].contiguous()
        c_loc = count[sl].contiguous()
        self.n_enq = int(count.max().item()) if count.numel() else 0
        self.ptr = int(ptr)
        self.qstate.copy_(torch.tensor([self.ptr, self.n_enq, 0, 0], dtype=torch.int64))
        self.max_key_norm = max(1.0, float(queue_ck.norm(dim=0).max().item()))
        _cabi.call("mscl_queue_import", self.queue.data_ptr(), self.queue_tf32.data_ptr(), self.birth.data_ptr(),
                   self.qstate.data_ptr(),
                   q_loc.data_ptr(), c_loc.data_ptr(), self.C, self.K_local, _stream())
        if self.world > 1:
            self.full_qstate.copy_(self.qstate)
            q_all = queue_ck.contiguous()
            _cabi.call("mscl_queue_import", self.full_queue.data_ptr(), None, self.full_birth.data_ptr(),
                       self.full_qstate.data_ptr(), q_all.data_ptr(), count.data_ptr(), self.C, self.K, _stream())
        self._fresh = True
        # keep the staging tensors alive until the kernel has consumed them
        torch.cuda.current_stream().synchronize()

    def export_full(self):
        """Return (queue_ck (C,K), count (K,) int64) of the WHOLE queue without a collective."""
        if self.world == 1:
            return self.export()
        q = torch.empty(self.C, self.K, device=self.device)
        c = torch.empty(self.K, dtype=torch.int64, device=self.device)
        _cabi.call("mscl_queue_export", self.full_queue.data_ptr(), self.full_birth.data_ptr(), self.full_qstate.data_ptr(),
                   q.data_ptr(), c.data_ptr(), self.C, self.K, _stream())
        return q, c

    def export(self):
        """Return (queue_ck (C,K_local), count (K_local,) int64) of this shard."""
        q = torch.empty(self.C, self.K_local, device=self.device)
        c = torch.empty(self.K_local, dtype=torch.int64, device=self.device)
        _cabi.call("mscl_queue_export", self.queue.data_ptr(), self.birth.data_ptr(), self.qstate.data_ptr(),
                   q.data_ptr(), c.data_ptr(), self.C, self.K_local, _stream())
        return q, c

    def weight(self):
        """Decayed snapshot (C, K_local) = queue * 0.99999**count (moco.py:484-486)."""
        w = torch.empty(self.C, self.K_local, device=self.device)
        _cabi.call("mscl_queue_weight", self.queue.data_ptr(), self.birth.data_ptr(), self.qstate.data_ptr(),
                   w.data_ptr(), self.C, self.K_local, _stream())
        return w

    @torch.no_grad()
    def enqueue(self, keys_all, save=False):
        """keys_all: (B_all, C) rank-major gathered keys (moco.py:423-440).  save=True (unsharded queues): returns
        (old_keys (B_all, C) in the tensor-core pass's operand form, old_birth int32 (B_all,), first slot) of the rows
        this enqueue overwrote -- what `infonce_multi` needs to score rows against the queue as it was before it."""
        _chk(keys_all, name="keys")
        b = keys_all.shape[0]
        if self.K % b != 0:
            raise AssertionError(f"K={self.K} must be a multiple of the gathered batch size {b}")  # moco.py:432
        if self.ptr + b > self.K:
            raise _cabi.MsclError("queue pointer is not aligned to the batch size (batch size changed mid-cycle)")
        saved = None
        if save:
            if self.world != 1:
                raise _cabi.MsclError("enqueue(save=True) needs an unsharded queue")
            saved = (torch.empty(b, self.C, device=self.device), torch.empty(b, dtype=torch.int32, device=self.device), self.ptr)
        _cabi.call("mscl_enqueue", self.queue.data_ptr(), self.queue_tf32.data_ptr(), self.birth.data_ptr(),
                   self.qstate.data_ptr(),
                   keys_all.data_ptr(), b, self.C, self.K, self.shard_begin, self.K_local,
                   saved[0].data_ptr() if save else None, saved[1].data_ptr() if save else None, _stream(),
                   algo_bytes=(3 if save else 2) * b * self.C * 4)
        if self.world > 1:      # the whole-queue fp32 master every rank keeps for a collective-free state_dict()
            _cabi.call("mscl_enqueue", self.full_queue.data_ptr(), None, self.full_birth.data_ptr(), self.full_qstate.data_ptr(),
                       keys_all.data_ptr(), b, self.C, self.K, 0, self.K, None, None, _stream())
        self.ptr = (self.ptr + b) % self.K
        self.n_enq += 1
        self._fresh = True
        return saved


# ----------------------------------------------------------------------------------------
# K1: fused InfoNCE
# ----------------------------------------------------------------------------------------
class PeerWorkspace:
    """Symmetric (same offset on every rank) device buffers mapped into every peer over NVLink, for the sharded
    queue's exchange: each rank's kernels STORE into their peers' buffers, a device-side barrier orders the phases.

        qp_all  [G * rows, 132]   the gathered packed queries: rank r's prep writes rows [r*M, (r+1)*M) on every rank
        acc     [G, rows, 132]    slot s = rank s's partial result for THIS rank's rows (written by rank s)

    Built on torch's symmetric memory (allocation, handle exchange, barrier); the data path is this repo's kernels.
    Creating one is collective over `group`."""

    def __init__(self, group, rows, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.group, self.rows = group, int(rows)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        n = self.world * self.rows * PACK_LD
        self.buf = symm_mem.empty(2 * n, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.qp_all, self.acc = self.buf[:n], self.buf[n:]
        self.qp_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.acc_ptrs = torch.tensor([p + 4 * n for p in ptrs], dtype=torch.int64, device=device)

    def barrier(self):
        self.hdl.barrier(channel=0)


def peer_workspace(nq, group, rows):
    """The queue's exchange workspace, (re)created collectively when more rows are needed."""
    ws = getattr(nq, "_peer_ws", None)
    if ws is None or ws.rows < rows or ws.group is not group:
        ws = PeerWorkspace(group, rows, nq.device)
        nq._peer_ws = ws
    return ws


EXCHANGE = "peer"      # "peer": NVLink peer stores from the kernels + device barriers; "nccl": all_gather / reduce_scatter


_FUSED_WS = {}


def _fused_workspace(dev, M, slot=0):
    """Zero-initialised statistics accumulator + CTA counter of the single-launch InfoNCE op (mscl_infonce_fused leaves it
    zero).  One per (device, stream, M, slot): two launches that may overlap -- or two jobs of one launch (slot) -- must
    not share it."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream(), int(M), int(slot))
    ws = _FUSED_WS.get(key)
    if ws is None:
        ws = _FUSED_WS[key] = torch.zeros(16 * M * 4 + 4, device=dev)
    return ws


def _prefetch_flag(nq):
    """MSCL_INFONCE_EARLY_PREFETCH unless the queue was (re)written since the last pass over it: a pass launched as a
    programmatic dependent of the launch before it may only prefetch queue tiles early when that launch did not write them."""
    early = 0 if getattr(nq, "_fresh", True) else 1
    nq._fresh = False
    return early


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, kpos, nq, rows_per_group, inv_T, impl, group, need_grad, dup_slot, dup_age):
        M = q.shape[0]
        dev = q.device
        st = _stream()
        world = dist.get_world_size(group) if (group is not None and nq.world > 1) else 1
        M_all = M * world
        n_groups = M // rows_per_group
        row_loss = torch.empty(2 * M, device=dev)
        group_out = torch.empty(n_groups, 4, device=dev)
        ctx.fused = impl == "fused" and world == 1
        if ctx.fused:
            # ONE launch: prep + tcgen05 pass + statistics reduction + finalize by the last CTA; the O slabs are summed
            # by the backward kernel
            n_part = _cabi.query("mscl_infonce_fused_parts", M, nq.K_local, sm_count(dev))
            ws = _fused_workspace(dev, M)
            part = torch.empty(n_part, M, PACK_LD, device=dev) if need_grad else None
            rowaux = torch.empty(M, 4, device=dev)
            _cabi.call("mscl_infonce_fused", q.data_ptr(), kpos.data_ptr(), M, nq.queue_tf32.data_ptr(), nq.birth.data_ptr(),
                       nq.qstate.data_ptr(), nq.K_local, inv_T, nq.max_key_norm,
                       dup_slot.data_ptr() if dup_slot is not None else None, dup_age, ws.data_ptr(),
                       part.data_ptr() if need_grad else None, n_part, rows_per_group,
                       int(need_grad), _prefetch_flag(nq), row_loss.data_ptr(), rowaux.data_ptr(), group_out.data_ptr(), st,
                       algo_bytes=infonce_algo_bytes(M, nq.K_local), algo_flops=(4 if need_grad else 2) * M * nq.K_local * DIM)
            if need_grad:
                ctx.save_for_backward(part, rowaux, kpos)
            ctx.rows_per_group = rows_per_group
            ctx.mark_non_differentiable(row_loss)
            return group_out, row_loss
        dq_unit = torch.empty(M, DIM, device=dev)
        qpack = torch.empty(M, PACK_LD, device=dev)
        fused_pass = impl in ("fused", "fused_pass")      # "fused_pass": the sharded-queue launch sequence on one rank (tests)
        dscale = None
        if not fused_pass:
            k_pad = (nq.K_local + 127) // 128 * 128
            dscale = torch.empty(k_pad, device=dev)
        peer = world > 1 and EXCHANGE == "peer" and impl != "simt"
        ws = peer_workspace(nq, group, M) if peer else None
        _cabi.call("mscl_infonce_prep", q.data_ptr(), kpos.data_ptr(), M, nq.birth.data_ptr(), nq.qstate.data_ptr(),
                   nq.K_local, inv_T, nq.max_key_norm, qpack.data_ptr(), dscale.data_ptr() if dscale is not None else None,
                   dup_slot.data_ptr() if dup_slot is not None else None, dup_age,
                   ws.qp_ptrs.data_ptr() if peer else None, world if peer else 0, ws.rank * M if peer else 0, st)
        if peer:                # every rank's rows were stored into every rank's table: wait for all of them
            ws.barrier()
            qpack_all = ws.qp_all[:M_all * PACK_LD].view(M_all, PACK_LD)
        elif world > 1:
            qpack_all = torch.empty(M_all, PACK_LD, device=dev)
            dist.all_gather_into_tensor(qpack_all, qpack, group=group)
        else:
            qpack_all = qpack
        if impl == "simt":      # validation twin: accumulates into one zeroed slab
            n_part = 1
            part = torch.zeros(1, M_all, PACK_LD, device=dev)
            _cabi.call("mscl_infonce_partial_simt", qpack_all.data_ptr(), M_all, nq.queue_tf32.data_ptr(), dscale.data_ptr(),
                       nq.K_local, nq.shard_begin, part.data_ptr(), int(need_grad), st)
        elif fused_pass:        # the pass of the single-launch form (64-key units, in-kernel decay scale), slabs out
            n_part = _cabi.query("mscl_infonce_fused_parts", M_all, nq.K_local, sm_count(dev))
            part = torch.empty(n_part, M_all, PACK_LD, device=dev)
            _cabi.call("mscl_infonce_pass", qpack_all.data_ptr(), M_all, nq.queue_tf32.data_ptr(), nq.birth.data_ptr(),
                       nq.qstate.data_ptr(), nq.K_local, nq.shard_begin, inv_T, part.data_ptr(), n_part, int(need_grad),
                       _prefetch_flag(nq), st, algo_bytes=infonce_algo_bytes(M_all, nq.K_local),
                       algo_flops=(4 if need_grad else 2) * M_all * nq.K_local * DIM)
        else:
            n_part = _cabi.query("mscl_infonce_num_partials", M_all, nq.K_local, sm_count(dev))
            part = torch.empty(n_part, M_all, PACK_LD, device=dev)
            _cabi.call("mscl_infonce_partial", qpack_all.data_ptr(), M_all, nq.queue_tf32.data_ptr(), dscale.data_ptr(),
                       nq.K_local, nq.shard_begin, part.data_ptr(), n_part, int(need_grad), st,
                       algo_bytes=infonce_algo_bytes(M_all, nq.K_local),
                       algo_flops=(4 if need_grad else 2) * M_all * nq.K_local * DIM)
        if peer:
            # the sum over this rank's CTA slabs goes straight into the row owners' accumulators (peer stores), then
            # every rank holds G slabs -- one per shard -- of its own rows, which finalize adds in rank order
            _cabi.call("mscl_infonce_reduce_scatter", part.data_ptr(), n_part, M_all, M, ws.acc_ptrs.data_ptr(), ws.rank, st,
                       algo_bytes=4 * PACK_LD * M_all * (n_part + 1))
            ws.barrier()
            part, n_part = ws.acc, world
        elif world > 1:         # local slabs -> one slab, summed across ranks, each rank keeps its own rows
            acc = torch.empty(M_all, PACK_LD, device=dev)
            _cabi.call("mscl_infonce_reduce", part.data_ptr(), n_part, M_all, acc.data_ptr(), st)
            part = torch.empty(1, M, PACK_LD, device=dev)
            dist.reduce_scatter_tensor(part.view(M, PACK_LD), acc, op=dist.ReduceOp.SUM, group=group)
            n_part = 1
        _cabi.call("mscl_infonce_finalize", qpack.data_ptr(), kpos.data_ptr(), part.data_ptr(), n_part, M, rows_per_group,
                   inv_T, int(need_grad), row_loss.data_ptr(), dq_unit.data_ptr(), group_out.data_ptr(), st)
        ctx.save_for_backward(dq_unit)
        ctx.rows_per_group = rows_per_group
        ctx.mark_non_differentiable(row_loss)
        return group_out, row_loss

    @staticmethod
    def backward(ctx, g_group, _g_rows):
        gout = g_group[:, 0].contiguous()
        if ctx.fused:
            part, rowaux, kpos = ctx.saved_tensors
            n_part, M = part.shape[0], part.shape[1]
            dq = torch.empty(M, DIM, device=part.device)
            _cabi.call("mscl_infonce_bwd_slabs", part.data_ptr(), n_part, M, kpos.data_ptr(), rowaux.data_ptr(), gout.data_ptr(),
                       ctx.rows_per_group, dq.data_ptr(), _stream(), algo_bytes=4 * M * (n_part * DIM + 2 * DIM + 4))
            return dq, None, None, None, None, None, None, None, None, None
        (dq_unit,) = ctx.saved_tensors
        M = dq_unit.shape[0]
        dq = torch.empty_like(dq_unit)
        _cabi.call("mscl_infonce_bwd", dq_unit.data_ptr(), gout.data_ptr(), M, ctx.rows_per_group, dq.data_ptr(), _stream())
        return dq, None, None, None, None, None, None, None, None, None


class _InfoNCEMulti(torch.autograd.Function):
    """Several independent fused InfoNCE terms in one launch (mscl_infonce_fused_multi); one backward kernel per job."""

    @staticmethod
    def forward(ctx, jobs, need_grad, split, *qs):
        import ctypes
        n = len(jobs)
        dev = qs[0].device
        st = _stream()
        Ms = [int(q.shape[0]) for q in qs]
        Ks = [int(j["nq"].K_local) for j in jobs]
        arr_i32 = lambda v: (ctypes.c_int32 * n)(*v)
        arr_i64 = lambda v: (ctypes.c_int64 * n)(*v)
        arr_f32 = lambda v: (ctypes.c_float * n)(*v)
        arr_ptr = lambda ts: (ctypes.c_void_p * n)(*[(t.data_ptr() if t is not None else None) for t in ts])
        n_part = _cabi.query("mscl_infonce_fused_parts_multi", n, arr_i32(Ms), arr_i64(Ks), sm_count(dev))
        ws = [_fused_workspace(dev, M, slot=i) for i, M in enumerate(Ms)]
        parts = [torch.empty(n_part, M, PACK_LD, device=dev) if need_grad else None for M in Ms]
        rowaux = [torch.empty(M, 4, device=dev) for M in Ms]
        row_loss = [torch.empty(2 * M, device=dev) for M in Ms]
        group_out = [torch.empty(M // j["rows_per_group"], 4, device=dev) for M, j in zip(Ms, jobs)]
        nbytes = sum(infonce_algo_bytes(M, K) for M, K in zip(Ms, Ks))
        flops = sum((4 if need_grad else 2) * M * K * DIM for M, K in zip(Ms, Ks))
        x_job, xkeys, xbirth, rep_begin, row_split = split if split is not None else (-1, None, None, 0, 0)
        _cabi.call("mscl_infonce_fused_multi_x", n, arr_ptr(qs), arr_ptr([j["kpos"] for j in jobs]), arr_i32(Ms),
                   arr_ptr([j["nq"].queue_tf32 for j in jobs]), arr_ptr([j["nq"].birth for j in jobs]),
                   arr_ptr([j["nq"].qstate for j in jobs]), arr_i64(Ks), arr_f32([j["inv_T"] for j in jobs]),
                   arr_f32([j["nq"].max_key_norm for j in jobs]), arr_ptr([j["dup_slot"] for j in jobs]),
                   arr_i32([j["dup_age"] for j in jobs]), arr_ptr(ws), arr_ptr(parts), n_part,
                   arr_i32([j["rows_per_group"] for j in jobs]), int(need_grad), arr_i32([_prefetch_flag(j["nq"]) for j in jobs]),
                   arr_ptr(row_loss), arr_ptr(rowaux), arr_ptr(group_out), x_job,
                   xkeys.data_ptr() if xkeys is not None else None, xbirth.data_ptr() if xbirth is not None else None,
                   rep_begin, xkeys.shape[0] if xkeys is not None else 0, row_split, st,
                   algo_bytes=nbytes + (xkeys.numel() * 4 if xkeys is not None else 0), algo_flops=flops)
        ctx.n = n
        ctx.need_grad = need_grad
        ctx.rpg = [j["rows_per_group"] for j in jobs]
        if need_grad:
            ctx.save_for_backward(*parts, *rowaux, *[j["kpos"] for j in jobs])
        ctx.mark_non_differentiable(*row_loss)
        out = []
        for g, r in zip(group_out, row_loss):
            out += [g, r]
        return tuple(out)

    @staticmethod
    def backward(ctx, *grads):
        n = ctx.n
        saved = ctx.saved_tensors
        parts, rowaux, kpos = saved[:n], saved[n:2 * n], saved[2 * n:]
        import ctypes
        live = [i for i in range(n) if grads[2 * i] is not None]
        dqs = [None] * n
        if live:        # one launch for the gradients of all jobs
            m = len(live)
            gouts = [grads[2 * i][:, 0].contiguous() for i in live]
            n_part = parts[live[0]].shape[0]
            Ms = [parts[i].shape[1] for i in live]
            for i in live:
                dqs[i] = torch.empty(parts[i].shape[1], DIM, device=parts[i].device)
            arr_ptr = lambda ts: (ctypes.c_void_p * m)(*[t.data_ptr() for t in ts])
            arr_i32 = lambda v: (ctypes.c_int32 * m)(*v)
            _cabi.call("mscl_infonce_bwd_slabs_multi", m, arr_ptr([parts[i] for i in live]), n_part, arr_i32(Ms),
                       arr_ptr([kpos[i] for i in live]), arr_ptr([rowaux[i] for i in live]), arr_ptr(gouts),
                       arr_i32([ctx.rpg[i] for i in live]), arr_ptr([dqs[i] for i in live]), _stream(),
                       algo_bytes=sum(4 * M * (n_part * DIM + 2 * DIM + 4) for M in Ms))
        return (None, None, None, *dqs)


def infonce_multi(jobs):
    """Several independent fused InfoNCE terms in ONE launch.  jobs: list (<= 4) of dicts with q, kpos (M,128), nq
    (an unsharded NegativeQueue), rows_per_group, T and optionally dup_slot / dup_age, as for `infonce`.  Returns a list
    of (group_out, row_stats) pairs.  The jobs occupy disjoint SMs and share the launch's fixed costs.

    ONE job may carry an "epoch split" (include/mscl_b200.h, mscl_infonce_fused_multi_x): `overwritten` = what
    `nq.enqueue(keys, save=True)` returned for the queue's LAST enqueue, and `row_split`: rows [0, row_split) are scored
    against the queue as it was before that enqueue (ages - 1, the B slots it wrote still holding the old keys), the rows
    from row_split on against the queue as it is."""
    if not 1 <= len(jobs) <= 4:
        raise _cabi.MsclError("infonce_multi takes 1 to 4 jobs")
    prepared, qs = [], []
    split = None
    for ji, j in enumerate(jobs):
        if j.get("overwritten") is not None:
            if split is not None:
                raise _cabi.MsclError("only one job of a launch may carry an epoch split")
            xk, xb, begin = j["overwritten"]
            _chk(xk, name="overwritten keys"), _chk(xb, torch.int32, "overwritten births")
            if xk.dim() != 2 or xk.shape[1] != DIM or not 1 <= xk.shape[0] <= 128 or xb.shape != (xk.shape[0],):
                raise _cabi.MsclError(f"overwritten keys must be (B <= 128, {DIM}) with B births; got {tuple(xk.shape)}")
            if (int(begin) + xk.shape[0]) % j["nq"].K != j["nq"].ptr:
                raise _cabi.MsclError("`overwritten` is not what the queue's last enqueue saved")
            split = (ji, xk, xb, int(begin), int(j["row_split"]))
        q, kpos, nq = j["q"], j["kpos"].detach(), j["nq"]
        _chk(q, name="q"), _chk(kpos, name="kpos")
        if q.dim() != 2 or q.shape[1] != DIM or kpos.shape != q.shape or q.shape[0] % j["rows_per_group"]:
            raise _cabi.MsclError("bad job shapes")
        if nq.world != 1:
            raise _cabi.MsclError("infonce_multi needs unsharded queues")
        dup = j.get("dup_slot")
        if dup is not None:
            _chk(dup, torch.int32, "dup_slot")
        prepared.append(dict(kpos=kpos, nq=nq, rows_per_group=int(j["rows_per_group"]), inv_T=float(1.0 / j["T"]),
                             dup_slot=dup, dup_age=int(j.get("dup_age", 1))))
        qs.append(q)
    need_grad = bool(torch.is_grad_enabled() and any(q.requires_grad for q in qs))
    flat = _InfoNCEMulti.apply(prepared, need_grad, split, *qs)
    return [(flat[2 * i], flat[2 * i + 1]) for i in range(len(jobs))]


def infonce_algo_bytes(M, K_local):
    """Algorithmic bytes of ONE fused pass: the queue shard once (fp32), the
    per-key decay scale, the packed queries in and the accumulator rows out."""
    return K_local * DIM * 4 + K_local * 4 + 2 * M * PACK_LD * 4


def infonce(q, kpos, nq, rows_per_group, T, impl="fused", group=None, dup_slot=None, dup_age=1):
    """Fused InfoNCE over the queue `nq`.

    q, kpos: (M, 128) stacked query rows and the positive key of each row; consecutive
    blocks of rows_per_group rows form one loss term (one "head call" of the reference).
    dup_slot: optional int32 (M,) GLOBAL queue slot that currently holds a copy of the row's own
    positive key (-1: none), dup_age its age -- see include/mscl_b200.h (K1, prep).
    impl: "fused" (default) -- one launch (csrc/infonce_fused.cu; with a sharded queue: its pass between the peer
    exchange kernels); "tc" -- the three-launch slab form (csrc/infonce_tc.cu, bit-reproducible); "simt" -- CUDA-core twin.
    Returns (group_out (M/rows_per_group, 4) = [loss, top1, top5, 0], row_stats (2M,)).
    Gradient flows to q only (keys and queue are detached in the reference, moco.py:486,532).
    """
    if dup_slot is not None:
        _chk(dup_slot, torch.int32, "dup_slot")
        if dup_slot.shape != (q.shape[0],):
            raise _cabi.MsclError("dup_slot must hold one slot per query row")
    _chk(q, name="q"), _chk(kpos, name="kpos")
    if q.dim() != 2 or q.shape[1] != DIM or kpos.shape != q.shape:
        raise _cabi.MsclError(f"q and kpos must both be (M, {DIM}); got {tuple(q.shape)} and {tuple(kpos.shape)}")
    if q.shape[0] % rows_per_group:
        raise _cabi.MsclError("number of rows must be a multiple of rows_per_group")
    need_grad = bool(q.requires_grad and torch.is_grad_enabled())
    return _InfoNCE.apply(q, kpos.detach(), nq, int(rows_per_group), float(1.0 / T), impl, group, need_grad,
                          dup_slot, int(dup_age))


# ----------------------------------------------------------------------------------------
# K2: LMCL
# ----------------------------------------------------------------------------------------
class _HWMean(torch.autograd.Function):
    """Mean over (H, W) of a dense (N,C,T,H,W) map -> (N,C,T) row-major.  Both dense layouts are read where they lie:
    row-major maps by the warp-per-row / staged-rows kernels, channels_last_3d maps (what the encoders produce in the
    channels-last step) by the NDHWC kernels; the input gradient comes back in the input's layout."""

    @staticmethod
    def forward(ctx, x):
        shape = x.shape
        HW = shape[-1] * shape[-2]
        R = x.numel() // HW
        out = torch.empty(shape[:-2], device=x.device)
        cl = x.dim() == 5 and _is_cl3d(x) and shape[1] % 32 == 0 and shape[0] * shape[2] <= 65535
        if cl:
            n, c, t = shape[:3]
            _cabi.call("mscl_hw_mean_ndhwc_fwd", x.data_ptr(), out.data_ptr(), n, c, t, HW, _stream(),
                       algo_bytes=4 * R * (HW + 1))
        else:
            x = x.contiguous()
            _cabi.call("mscl_hw_mean_fwd", x.data_ptr(), out.data_ptr(), R, HW, _stream(), algo_bytes=4 * R * (HW + 1))
        ctx.shape = shape
        ctx.cl = cl
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        shape = ctx.shape
        HW = shape[-1] * shape[-2]
        if ctx.cl:
            n, c, t = shape[:3]
            gx = torch.empty(shape, device=g.device, memory_format=torch.channels_last_3d)
            _cabi.call("mscl_hw_mean_ndhwc_bwd", g.data_ptr(), gx.data_ptr(), n, c, t, HW, _stream(),
                       algo_bytes=4 * g.numel() * (HW + 1))
        else:
            gx = torch.empty(shape, device=g.device)
            _cabi.call("mscl_hw_mean_bwd", g.data_ptr(), gx.data_ptr(), g.numel(), HW, _stream(),
                       algo_bytes=4 * g.numel() * (HW + 1))
        return gx


def hw_mean(x):
    """Mean over the last two dims: AdaptiveAvgPool3d((None,1,1)).view(b,c,t) (local_cl_head.py:61-62) for a dense fp32
    CUDA map, row-major or channels_last_3d (any other stride pattern is made row-major first)."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise _cabi.MsclError("feature map must be a CUDA tensor (the MSCL hot path has no CPU fallback)")
    if x.dtype != torch.float32:
        raise _cabi.MsclError(f"feature map must be torch.float32, got {x.dtype}")
    _cabi.require_device(x.device.index if x.device.index is not None else torch.cuda.current_device())
    if not (x.is_contiguous() or (x.dim() == 5 and _is_cl3d(x))):
        x = x.contiguous()
    return _HWMean.apply(x)


_LMCL_PART = {}


class _LMCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xq, xf, inv_T):
        N, C, t = xq.shape
        t2 = xf.shape[2]
        dev = xq.device
        key = (dev.index, N)
        if key not in _LMCL_PART:
            _LMCL_PART[key] = torch.zeros(N * 4 + 4, device=dev)
        part = _LMCL_PART[key]
        out = torch.empty(4, device=dev)
        gxq = torch.empty_like(xq)
        gxf = torch.empty_like(xf)
        _cabi.call("mscl_lmcl", xq.data_ptr(), xf.data_ptr(), N, C, t, t2, inv_T, out.data_ptr(), gxq.data_ptr(),
                   gxf.data_ptr(), part.data_ptr(), _stream(), algo_bytes=8 * N * C * (t + t2),
                   algo_flops=6 * N * t * t2 * C)
        ctx.save_for_backward(gxq, gxf)
        return out

    @staticmethod
    def backward(ctx, g):
        gxq, gxf = ctx.saved_tensors
        s = g[0]
        return gxq * s, gxf * s, None


def lmcl(xq, xf, T):
    """LMCL loss on pooled features xq (N,C,t), xf (N,C,2t) -> tensor [loss, top1, top5, 0]
    (local_cl_head.py:63-73 + :41-51)."""
    _chk(xq, name="xq"), _chk(xf, name="xf")
    if xq.dim() != 3 or xf.dim() != 3 or xq.shape[:2] != xf.shape[:2] or xf.shape[2] < xq.shape[2]:
        raise _cabi.MsclError(f"bad LMCL shapes {tuple(xq.shape)} {tuple(xf.shape)}")
    return _LMCL.apply(xq, xf, float(1.0 / T))


# ----------------------------------------------------------------------------------------
# K4: multi-tensor EMA
# ----------------------------------------------------------------------------------------
class EmaTable:
    """Pointer tables for one launch over every (param_k, param_q) pair (moco.py:416-421)."""

    CHUNK = 16384

    def __init__(self, params_k, params_q):
        self.params_k = [p for p in params_k]
        self.params_q = [p for p in params_q]
        if len(self.params_k) != len(self.params_q):
            raise _cabi.MsclError("key/query parameter lists differ in length")
        self._sig = None
        self._build()

    def _signature(self):
        return tuple(p.data_ptr() for p in self.params_k) + tuple(p.data_ptr() for p in self.params_q)

    def _build(self):
        dev = self.params_k[0].device
        for pk, pq in zip(self.params_k, self.params_q):
            _chk_dense(pk.data, "param_k"), _chk_dense(pq.data, "param_q")
            # same memory ORDER is what the element-wise walk needs: literal strides may differ where an extent is 1
            # (a [32,16,1,1,1] weight is row-major with strides (16,1,1,1,1) and with (16,1,16,16,16) alike)
            same_order = pk.stride() == pq.stride() or (pk.is_contiguous() and pq.is_contiguous())
            if pk.shape != pq.shape or not same_order:
                raise _cabi.MsclError("key/query parameter shapes or memory layouts differ")
        sizes = [p.numel() for p in self.params_k]
        blk_t, blk_s = [], []
        for i, n in enumerate(sizes):
            for s in range(0, n, self.CHUNK):
                blk_t.append(i)
                blk_s.append(s)
        self.n_blocks = len(blk_t)
        self.numel = sum(sizes)
        self.k_ptrs = torch.tensor([p.data_ptr() for p in self.params_k], dtype=torch.int64, device=dev)
        self.q_ptrs = torch.tensor([p.data_ptr() for p in self.params_q], dtype=torch.int64, device=dev)
        self.sizes = torch.tensor(sizes, dtype=torch.int64, device=dev)
        self.blk_tensor = torch.tensor(blk_t, dtype=torch.int32, device=dev)
        self.blk_start = torch.tensor(blk_s, dtype=torch.int64, device=dev)
        self._sig = self._signature()

    @torch.no_grad()
    def update(self, m):
        """k <- k*m + q*(1-m) with m a python float, rounded like the reference's scalar multiply."""
        if self._signature() != self._sig:
            self._build()      # parameters were re-allocated (.cuda(), load_state_dict with assign, ...)
        m32 = float(np.float32(m))
        om32 = float(np.float32(1.0 - m))
        _cabi.call("mscl_ema_multi", self.k_ptrs.data_ptr(), self.q_ptrs.data_ptr(), self.sizes.data_ptr(),
                   self.blk_tensor.data_ptr(), self.blk_start.data_ptr(), self.n_blocks, self.CHUNK, m32, om32, _stream(),
                   algo_bytes=12 * self.numel, algo_flops=3 * self.numel)


# ----------------------------------------------------------------------------------------
# K3: FRA
# ----------------------------------------------------------------------------------------
def fra_table(ratios=(0.2, 1.8), num_chunks=8, device="cuda"):
    """(cos, sin) float32 pairs of beta = (start + stride*cid)*pi (transforms_motion.py:106-124)."""
    start = ratios[0]
    stride = (ratios[1] - ratios[0]) / num_chunks
    tab = [[math.cos((start + stride * c) * math.pi), math.sin((start + stride * c) * math.pi)]
           for c in range(num_chunks)]
    return torch.tensor(np.array(tab, dtype=np.float64).astype(np.float32), device=device).contiguous()


FRA_FUSED_MAX_PIXELS = 204800


@torch.no_grad()
def fra(flow, cid, table, layout="planar", one_pass=None):
    """Normalised base + rotated flow.  flow: planar (N,2,T,H,W) or interleaved (N,T,H,W,2);
    cid int32 (N,).  Returns (N,2,2T,H,W): base frames then FRA frames (transforms_motion.py:111-142)."""
    _chk(flow, name="flow"), _chk(cid, torch.int32, "cid"), _chk(table, name="table")
    if layout == "planar":
        N, two, T, H, W = flow.shape
        lay = 0
    else:
        N, T, H, W, two = flow.shape
        lay = 1
    if two != 2:
        raise _cabi.MsclError("flow must have exactly two components (u, v)")
    out = torch.empty(N, 2, 2 * T, H, W, device=flow.device)
    st = _stream()
    if one_pass is None:
        one_pass = H * W <= FRA_FUSED_MAX_PIXELS
    if one_pass:                           # one pass: a cluster of 8 CTAs keeps the frame in shared memory
        _cabi.call("mscl_fra_fused", flow.data_ptr(), cid.data_ptr(), table.data_ptr(), out.data_ptr(), N, T, H * W, lay, st,
                   algo_bytes=24 * N * T * H * W)
        return out
    maxrad = torch.empty(N, T, 2, device=flow.device)
    _cabi.call("mscl_fra_maxrad", flow.data_ptr(), cid.data_ptr(), table.data_ptr(), maxrad.data_ptr(), N, T, H * W, lay, st,
               algo_bytes=8 * N * T * H * W)
    _cabi.call("mscl_fra_apply", flow.data_ptr(), cid.data_ptr(), table.data_ptr(), maxrad.data_ptr(), out.data_ptr(),
               N, T, H * W, lay, st, algo_bytes=24 * N * T * H * W)
    return out


_cabi._LAUNCHES_PER_CALL["mscl_fra_maxrad"] = 2  # memset + kernel


@torch.no_grad()
def fra_rotate(flow, cid, table):
    """Rotation only of an already normalised planar clip (N,2,T,H,W)."""
    _chk(flow, name="flow"), _chk(cid, torch.int32, "cid"), _chk(table, name="table")
    N, two, T, H, W = flow.shape
    out = torch.empty_like(flow)
    _cabi.call("mscl_fra_rotate", flow.data_ptr(), cid.data_ptr(), table.data_ptr(), out.data_ptr(), N, T, H * W, _stream(),
               algo_bytes=16 * N * T * H * W)
    return out


# ----------------------------------------------------------------------------------------
# K6: row gather for shuffle-BN
# ----------------------------------------------------------------------------------------
@torch.no_grad()
def gather_rows(x, idx):
    """out[r] = x[idx[r]] along dim 0 (moco.py:172,191)."""
    _chk(x, name="x"), _chk(idx, torch.int64, "idx")
    row = x[0].numel()
    out = torch.empty((idx.numel(),) + tuple(x.shape[1:]), device=x.device)
    _cabi.call("mscl_gather_rows", x.data_ptr(), idx.data_ptr(), out.data_ptr(), idx.numel(), row, _stream(),
               algo_bytes=8 * idx.numel() * row)
    return out


# ----------------------------------------------------------------------------------------
# K7: trilinear up-sampling (TPN neck)
# ----------------------------------------------------------------------------------------
class _UpsampleTrilinear(torch.autograd.Function):
    """Both dense layouts are served natively: a channels_last_3d input gives a channels_last_3d output (and gradient),
    so the TPN neck's `y + upsample(...)` stays a same-layout add and no NCDHW <-> NDHWC copy appears on either side."""

    @staticmethod
    def forward(ctx, x, size):
        n, c, ti, hi, wi = x.shape
        to, ho, wo = size
        cl = _is_cl3d(x) and c % 4 == 0
        if cl:
            y = torch.empty((n, c, to, ho, wo), device=x.device, memory_format=torch.channels_last_3d)
            _cabi.call("mscl_upsample_trilinear_ndhwc_fwd", x.data_ptr(), y.data_ptr(), n, c, ti, hi, wi, to, ho, wo,
                       _stream(), algo_bytes=4 * (x.numel() + y.numel()))
        else:
            x = x.contiguous()
            y = torch.empty(n, c, to, ho, wo, device=x.device)
            _cabi.call("mscl_upsample_trilinear_fwd", x.data_ptr(), y.data_ptr(), n * c, ti, hi, wi, to, ho, wo, _stream(),
                       algo_bytes=4 * (x.numel() + y.numel()))
        ctx.in_shape = tuple(x.shape)
        ctx.cl = cl
        return y

    @staticmethod
    def backward(ctx, gy):
        n, c, ti, hi, wi = ctx.in_shape
        to, ho, wo = gy.shape[2:]
        if ctx.cl:
            # separable: one 1-D gather pass per scaled axis (W, H, T), each reading its input once
            cur = gy.contiguous(memory_format=torch.channels_last_3d)
            c4 = c // 4
            dims = [to, ho, wo]                                   # current (T, H, W) extent of `cur`
            for axis, (n_in, n_out) in ((2, (wi, wo)), (1, (hi, ho)), (0, (ti, to))):
                if n_in == n_out:
                    continue
                outer = n * int(np.prod(dims[:axis], dtype=np.int64))
                inner4 = int(np.prod(dims[axis + 1:], dtype=np.int64)) * c4
                dims[axis] = n_in
                nxt = torch.empty((n, c, dims[0], dims[1], dims[2]), device=gy.device, memory_format=torch.channels_last_3d)
                _cabi.call("mscl_linear_axis_bwd", cur.data_ptr(), nxt.data_ptr(), outer, n_in, n_out, inner4, _stream(),
                           algo_bytes=4 * (cur.numel() + nxt.numel()))
                cur = nxt
            gx = cur if cur.data_ptr() != gy.data_ptr() else cur.clone()
        else:
            gy = gy.contiguous()
            gx = torch.empty(ctx.in_shape, device=gy.device)
            _cabi.call("mscl_upsample_trilinear_bwd", gy.data_ptr(), gx.data_ptr(), n * c, ti, hi, wi, to, ho, wo, _stream(),
                       algo_bytes=4 * (gx.numel() + gy.numel()))
        return gx, None


def upsample_trilinear(x, size):
    """F.interpolate(x, size=size, mode="trilinear") with align_corners=False (necks/sepc.py:126-130) for a dense
    fp32 (N,C,T,H,W) CUDA tensor, row-major or channels_last_3d (any other stride pattern is made row-major first)."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.float32:
        raise _cabi.MsclError("upsample_trilinear needs a float32 CUDA tensor (no CPU fallback)")
    if x.dim() != 5 or len(size) != 3:
        raise _cabi.MsclError("upsample_trilinear takes (N,C,T,H,W) and a (T,H,W) size")
    _cabi.require_device(x.device.index if x.device.index is not None else torch.cuda.current_device())
    if not (_is_cl3d(x) or x.is_contiguous()):
        x = x.contiguous()
    return _UpsampleTrilinear.apply(x, tuple(int(v) for v in size))


# ----------------------------------------------------------------------------------------
# K8 / K9: augmentation front-end
# ----------------------------------------------------------------------------------------
@torch.no_grad()
def flow_visualize(flow, flip=None, norm=None):
    """Colour-wheel image of a planar flow clip (N,2,T,H,W) -> (N,3,T,H,W) (ssl_aug.py:87-136), mirrored along W
    for the samples of the uint8 mask `flip` (ssl_aug_v2.py:109-117), optionally normalised by norm = [mean3, std3]."""
    _chk(flow, name="flow")
    if flow.dim() != 5 or flow.shape[1] != 2:
        raise _cabi.MsclError("flow must be (N,2,T,H,W)")
    N, _, T, H, W = flow.shape
    if flip is not None:
        _chk(flip, torch.uint8, "flip")
    if norm is not None:
        _chk(norm, name="norm")
    out = torch.empty(N, 3, T, H, W, device=flow.device)
    _cabi.call("mscl_flow_visualize", flow.data_ptr(), flip.data_ptr() if flip is not None else None,
               norm.data_ptr() if norm is not None else None, out.data_ptr(), N, T, H, W, _stream(),
               algo_bytes=20 * N * T * H * W)
    return out


COLOR_PARAMS = 16
_GRAY_CHUNKS = 16


@torch.no_grad()
def color_pipeline(x, params, taps, norm):
    """Fused flip / colour jitter / grayscale / Gaussian blur / normalise of RGB clips (N,3,T,H,W); params (N,16): one
    set per clip, or (N*T,16): one set per frame (row n*T + t), as described in include/mscl_b200.h (K9); taps the
    odd-length 1-D blur kernel, norm = [mean3, std3]."""
    _chk(x, name="clips"), _chk(params, name="params"), _chk(taps, name="taps"), _chk(norm, name="norm")
    if x.dim() != 5 or x.shape[1] != 3 or params.dim() != 2 or params.shape[1] != COLOR_PARAMS:
        raise _cabi.MsclError("clips must be (N,3,T,H,W) and params (N,16) or (N*T,16)")
    N, _, T, H, W = x.shape
    if params.shape[0] not in (N, N * T):
        raise _cabi.MsclError(f"params must have {N} (per clip) or {N * T} (per frame) rows, got {params.shape[0]}")
    per_frame = int(params.shape[0] == N * T and T > 1)
    out = torch.empty_like(x)
    chunks = 1 if per_frame else _GRAY_CHUNKS
    scratch = torch.empty(params.shape[0], chunks, device=x.device)
    _cabi.call("mscl_color_pipeline", x.data_ptr(), params.data_ptr(), taps.data_ptr(), taps.numel(), norm.data_ptr(),
               scratch.data_ptr(), chunks, out.data_ptr(), N, T, H, W, per_frame, _stream(),
               algo_bytes=(24 if per_frame else 36) * N * T * H * W)      # per-clip factors: the luminance pre-pass reads x again
    return out
