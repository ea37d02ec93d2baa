import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
for count in (32 * 264, 8):
    x = torch.ones(count, dtype=torch.float64, device="cuda")
    out = torch.empty(count * world, dtype=torch.float64, device="cuda")
    for _ in range(20):
        dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        dist.all_gather_into_tensor(out, x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("world %d all_gather of %d doubles per rank: %.1f us per call" % (world, count, e0.elapsed_time(e1) * 1e3 / 200), flush=True)
dist.destroy_process_group()
