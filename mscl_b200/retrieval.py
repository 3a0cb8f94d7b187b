"""Nearest-neighbour retrieval evaluation on the device (reference: tools/test_retrival.py:286-304).

    accs = nn_retrieval_accuracy(train_feature, test_feature, train_label, test_label)   # ks = (1, 5, 10, 20, 50)

Same arithmetic as the reference's script -- centre both feature sets by their own column mean, L2-normalise the rows,
`sim = test @ train.T`, acc@k = fraction of test items with a same-label item among their k nearest train items -- with
the streaming ends in two kernels (K11, csrc/retrieval.cu): `center_normalize`, and `retrieval_rank`, which turns the
five top-k + label-gather rounds into one pass over `sim`.  The GEMM is a plain fp32 library GEMM (cuBLAS).
CUDA tensors only: a host tensor raises MsclError.
"""
import torch

from . import _cabi
from .functional import _chk, _stream

KS = (1, 5, 10, 20, 50)
_CHUNKS = 64


@torch.no_grad()
def center_normalize(x):
    """x (N, D) fp32 -> (x - x.mean(0)) rows scaled to unit L2 norm (eps 1e-12)  (test_retrival.py:290-296)."""
    _chk(x, name="features")
    if x.dim() != 2:
        raise _cabi.MsclError("features must be (N, D)")
    N, D = x.shape
    chunks = max(1, min(_CHUNKS, N))
    partial = torch.empty(chunks, D, dtype=torch.float64, device=x.device)
    mean = torch.empty(D, device=x.device)
    out = torch.empty_like(x)
    _cabi.call("mscl_center_normalize", x.data_ptr(), N, D, partial.data_ptr(), chunks, mean.data_ptr(), out.data_ptr(),
               _stream(), algo_bytes=12 * N * D)
    return out


@torch.no_grad()
def retrieval_rank(sim, train_label, test_label):
    """rank (n_test,) int32: how many train items score strictly above the test item's best same-label train item."""
    _chk(sim, name="sim"), _chk(train_label, torch.int64, "train_label"), _chk(test_label, torch.int64, "test_label")
    if sim.dim() != 2 or tuple(sim.shape) != (test_label.numel(), train_label.numel()):
        raise _cabi.MsclError(f"sim must be (n_test, n_train), got {tuple(sim.shape)}")
    n_test, n_train = sim.shape
    rank = torch.empty(n_test, dtype=torch.int32, device=sim.device)
    _cabi.call("mscl_retrieval_rank", sim.data_ptr(), sim.stride(0), train_label.data_ptr(), test_label.data_ptr(), n_test,
               n_train, rank.data_ptr(), _stream(), algo_bytes=4 * n_test * n_train)
    return rank


@torch.no_grad()
def nn_retrieval_accuracy(train_feature, test_feature, train_label, test_label, ks=KS):
    """kNN retrieval accuracies [acc@k for k in ks] as python floats (test_retrival.py:286-304)."""
    assert len(train_feature) == len(train_label), f"{len(train_feature)} vs {len(train_label)}"
    assert len(test_feature) == len(test_label), f"{len(test_feature)} vs {len(test_label)}"
    test = center_normalize(test_feature.contiguous().float())
    train = center_normalize(train_feature.contiguous().float())
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False          # the reference multiplies in fp32 on the host
    try:
        sim = test.matmul(train.t())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    rank = retrieval_rank(sim.contiguous(), train_label.contiguous().long(), test_label.contiguous().long())
    kt = torch.tensor(list(ks), dtype=torch.int32, device=rank.device)
    return (rank.view(-1, 1) < kt.view(1, -1)).float().mean(dim=0).tolist()       # one D2H copy
