"""GPU debug: compare the raw accumulators of the InfoNCE passes with a float64 torch model."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from mscl_b200 import functional as fx, _cabi
from test_gpu_kernels import _make_case

M, K = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 4096
q, kpos, queue, count = _make_case(M + K, M, K, 8)
nq = fx.NegativeQueue(K)
nq.load(queue, count, 0)
qd, kd = q.cuda(), kpos.cuda()
inv_T = 1 / 0.07
st = torch.cuda.current_stream().cuda_stream
for impl in ("simt", "tc"):
    qpack = torch.empty(M, 132, device="cuda")
    dscale = torch.empty((K + 63) // 64 * 64, device="cuda")
    acc = torch.empty(M, 132, device="cuda")
    _cabi.call("mscl_infonce_prep", qd.data_ptr(), kd.data_ptr(), M, nq.birth.data_ptr(), nq.qstate.data_ptr(), K, inv_T, 1.0,
               qpack.data_ptr(), dscale.data_ptr(), acc.data_ptr(), M, st)
    if impl == "simt":
        _cabi.call("mscl_infonce_partial_simt", qpack.data_ptr(), M, nq.queue_tf32.data_ptr(), dscale.data_ptr(), K, acc.data_ptr(), 1, st)
    else:
        _cabi.call("mscl_infonce_partial", qpack.data_ptr(), M, nq.queue_tf32.data_ptr(), dscale.data_ptr(), K, acc.data_ptr(), 1, 148, st)
    torch.cuda.synchronize()
    # float64 model of the accumulators
    qp = qpack[:, :128].double()
    ds = dscale[:K].double()
    W = nq.queue_tf32.double()                        # (K, C)
    s2 = (qp @ W.T) * ds                         # (M, K)
    shift2 = qpack[:, 129].double().unsqueeze(1)
    pos2 = qpack[:, 128].double().unsqueeze(1)
    p = torch.exp2(s2 - shift2)
    O = (p * ds) @ W
    ssum = p.sum(1)
    cnt = (s2 > pos2).sum(1).double()
    a = acc.double()
    print(impl, "O rel err", float((a[:, :128] - O).norm() / O.norm()), "sum rel", float(((a[:, 128] - ssum) / ssum).abs().max()),
          "cnt diff", float((a[:, 129] - cnt).abs().max()), "|O|", float(O.norm()), "|acc O|", float(a[:, :128].norm()))
    if M <= 8:
        print(" ratio O[0,:6]", (a[0, :6] / O[0, :6]).tolist())
    # finalize
    rpg = M
    row_loss = torch.empty(2 * M, device="cuda"); dq_unit = torch.empty(M, 128, device="cuda"); gout = torch.empty(1, 4, device="cuda")
    _cabi.call("mscl_infonce_finalize", qpack.data_ptr(), kd.data_ptr(), acc.data_ptr(), M, rpg, inv_T, row_loss.data_ptr(), dq_unit.data_ptr(), gout.data_ptr(), st)
    torch.cuda.synchronize()
    e0 = torch.exp2(pos2 - shift2).squeeze(1)
    Z = e0 + a[:, 128]
    p0 = e0 / Z
    g = ((p0 - 1).unsqueeze(1) * kd.double() * inv_T + a[:, :128] / Z.unsqueeze(1) * 0.6931471805599453) / rpg
    print(impl, "finalize dq_unit rel err vs formula-on-acc", float((dq_unit.double() - g).norm() / g.norm()),
          "p0", p0[:4].tolist(), "loss", float(gout[0, 0]), "model loss", float(((shift2.squeeze(1) + torch.log2(Z) - pos2.squeeze(1)) * 0.6931471805599453).mean()))
