"""Feasibility probe: does torch's symmetric memory rendezvous work on this box (peer pointers over NVLink)?"""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok", type(h).__name__, [hex(p) for p in h.buffer_ptrs][:4], "signal pads", len(h.signal_pad_ptrs), flush=True)
    t.fill_(rank + 1)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.uint8)
    print(rank, "peer first bytes", peer[:4].tolist(), flush=True)
    h.barrier()
    # latency of a tiny NCCL all_gather for comparison
    x = torch.zeros(8192, dtype=torch.int32, device=dev); out = torch.empty(world * 8192, dtype=torch.int32, device=dev)
    for _ in range(20): dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): dist.all_gather_into_tensor(out, x)
    b.record(); b.synchronize()
    print(rank, "NCCL all_gather 32 KB: %.1f us each (back to back)" % (a.elapsed_time(b) * 10), flush=True)
except Exception as e:
    print(rank, "symmetric memory unavailable:", repr(e)[:300], flush=True)
dist.barrier(); dist.destroy_process_group()
