"""GPU (2 ranks): does torch's symmetric memory give peer pointers + a device barrier on this box?"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
t = symm_mem.empty(1024, dtype=torch.float32, device=torch.device("cuda", rank))
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs], flush=True)
t.fill_(rank + 1)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
hdl.barrier()
peer[1] = 100 + rank          # remote store
hdl.barrier()
torch.cuda.synchronize()
print(rank, "local[1] written by peer:", float(t[1]), "multicast_ptr", getattr(hdl, "multicast_ptr", None), flush=True)
dist.destroy_process_group()
