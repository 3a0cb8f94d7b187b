"""1-GPU probe: the inputs of tests/test_gpu_multi.py at world 8 through the REPLICATED path, per 8-row rank slice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_step import head_level_model
from oracle import inputs, mscl_oracle as O
for world, N, K in ((8, 8, 4096), (4, 8, 4096), (8, 32, 65536)):
    t = 4
    inp = inputs.head_inputs(seed=11, N=N * world, K=K, t=t, hw_rgb=6, hw_flow=3, b_all=N * world)
    model = head_level_model(K, t)
    model.train()
    ptr = torch.tensor([inp["ptr"]])
    model.load_state_dict({"recognizer.queue": inp["queue_rgb"], "recognizer.count": inp["count"], "recognizer.queue_ptr": ptr,
                           "recognizer_flow.queue": inp["queue_flow"], "recognizer_flow.count": inp["count"],
                           "recognizer_flow.queue_ptr": ptr}, strict=False)
    names = ("q", "q_f", "q_af", "q_map", "qf_map", "qaf_map")
    leaves = {n: inp[n].cuda().requires_grad_(True) for n in names}
    feats = dict(q=leaves["q"], q_f=leaves["q_f"], q_af=leaves["q_af"], k=inp["k"].cuda(), k_f=inp["k_f"].cuda(),
                 k_af=inp["k_af"].cuda(), q_mlvl=[leaves["q_map"]], q_flow_mlvl=[leaves["qf_map"]], q_aug_flow_mlvl=[leaves["qaf_map"]])
    losses = model.objective(feats)
    sum(v.mean() for k, v in losses.items() if "loss" in k).backward()
    ol = {n: inp[n].clone().requires_grad_(True) for n in names}
    of = dict(k=inp["k"], k_f=inp["k_f"], k_af=inp["k_af"], **ol)
    ref = O.mscl_objective(of, O.QueueState(inp["queue_rgb"], inp["count"], inp["ptr"]), O.QueueState(inp["queue_flow"], inp["count"], inp["ptr"]), T=0.07, t=t)
    sum(v.mean() for k, v in ref.items() if "loss" in k).backward()
    for n in ("q", "q_f", "q_af"):
        a, b = leaves[n].grad.cpu().double(), ol[n].grad.double()
        per = [float((a[r * N:(r + 1) * N] - b[r * N:(r + 1) * N]).norm() / b[r * N:(r + 1) * N].norm()) for r in range(world)]
        rows = ((a - b).norm(dim=1) / b.norm(dim=1))
        print(f"world {world} N={N} K={K} grad {n}: whole {float((a - b).norm() / b.norm()):.2e}  per-rank-slice max {max(per):.2e}  worst row {float(rows.max()):.2e} (|g| of that row {float(b.norm(dim=1)[rows.argmax()]):.2e}, median |g| {float(b.norm(dim=1).median()):.2e})")
    for k in losses:
        if "acc" in k:
            print("   ", k, float(losses[k].mean()), float(ref[k].mean()))
