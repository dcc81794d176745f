"""Developer probe (torchrun, N ranks): NCCL transport and all-gather latency for the sizes bench.py exchanges per step."""
import os
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    if rank == 0:
        print("peer access 0->1:", torch.cuda.can_device_access_peer(0, 1), flush=True)
    for name, numel in (("desc 16x2000x128 f32", 16 * 2000 * 128), ("lafs 16x2000x6 f32", 16 * 2000 * 6), ("counts 16 i32", 16)):
        src = torch.ones(numel, device=dev)
        dst = torch.empty(world * numel, device=dev)
        for _ in range(5):
            dist.all_gather_into_tensor(dst, src)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_gather_into_tensor(dst, src)
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print("%-24s %8.3f ms per all-gather (%.1f MB per rank)" % (name, e0.elapsed_time(e1) / 20, numel * 4 / 1e6), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
