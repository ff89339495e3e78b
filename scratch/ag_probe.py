import os, torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
dev = torch.device('cuda', int(os.environ['LOCAL_RANK'])); torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
for per_rows in (2449029 // world + 1,):
    x = torch.randn(per_rows, 128, device=dev)
    full = torch.empty(per_rows * world, 128, device=dev)
    for _ in range(3): dist.all_gather_into_tensor(full, x)
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): dist.all_gather_into_tensor(full, x)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    recv = x.numel() * 4 * (world - 1)
    if rank == 0:
        print(f'all_gather world={world} out={full.numel()*4/1e6:.0f}MB: {ms:.3f} ms, recv {recv/ms/1e6:.0f} GB/s per rank, busbw {full.numel()*4*(world-1)/world/ms/1e6:.0f} GB/s')
dist.destroy_process_group()
