import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1024, dtype=torch.int64, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "symm ok", type(hdl).__name__, [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
    t.fill_(rank + 1)
    dist.barrier(); torch.cuda.synchronize()
    peer = hdl.get_buffer((rank + 1) % world, (8,), torch.int64)
    print(rank, "peer view", peer.tolist(), flush=True)
    print(rank, "attrs", [a for a in dir(hdl) if not a.startswith("_")], flush=True)
except Exception as e:
    print(rank, "symm FAILED", repr(e), flush=True)
dist.barrier()
dist.destroy_process_group()
