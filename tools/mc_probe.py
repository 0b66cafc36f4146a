"""Probe: torch symmetric memory + NVSwitch multicast on this box (run under torch.distributed.run)."""
import ctypes, os, sys, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N = 64 << 20
t = symm_mem.empty(N, dtype=torch.uint8, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "mc", hex(hdl.multicast_ptr or 0), "has_mc", hdl.has_multicast_support(torch._C._distributed_c10d.DeviceType.CUDA if hasattr(torch._C._distributed_c10d, "DeviceType") else "cuda", lr) if False else "", flush=True)
mc = hdl.multicast_ptr
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
cudart.cudaMemcpyAsync.restype = ctypes.c_int

class Raw:
    def __init__(self, ptr, n): self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}

t.zero_(); torch.cuda.synchronize(); dist.barrier()
chunk = 4 << 20
src = torch.full((chunk,), rank + 1, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
if mc:
    # 1) memcpy to the multicast address: every rank writes its chunk at offset rank * chunk
    rc = cudart.cudaMemcpyAsync(mc + rank * chunk, src.data_ptr(), chunk, 4, st)
    torch.cuda.synchronize(); dist.barrier()
    got = [int(t[q * chunk].item()) for q in range(world)] + [int(t[q * chunk + chunk - 1].item()) for q in range(world)]
    print(rank, "memcpy->mc rc", rc, "local view after all ranks wrote:", got, flush=True)
    # timing
    dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): cudart.cudaMemcpyAsync(mc + rank * chunk, src.data_ptr(), chunk, 4, st)
    e1.record(); torch.cuda.synchronize()
    print(rank, f"memcpy->mc 4 MiB: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us each", flush=True)
    dist.barrier()
    # 2) kernel stores to the multicast address (torch fill_ kernel on a tensor view of the mc pointer)
    try:
        mct = torch.as_tensor(Raw(mc + (8 + rank) * chunk, chunk), device=dev)
        mct.fill_(10 + rank)
        torch.cuda.synchronize(); dist.barrier()
        print(rank, "kernel store->mc:", [int(t[(8 + q) * chunk + 5].item()) for q in range(world)], flush=True)
        e0.record()
        for _ in range(20): mct.fill_(3)
        e1.record(); torch.cuda.synchronize()
        print(rank, f"fill kernel->mc 4 MiB: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us each", flush=True)
    except Exception as e:
        print(rank, "kernel store to mc failed:", repr(e), flush=True)
    # 3) 4-byte flag copy to mc
    dist.barrier()
    e0.record()
    for _ in range(20): cudart.cudaMemcpyAsync(mc + 15 * chunk + 64 * rank, src.data_ptr(), 4, 4, st)
    e1.record(); torch.cuda.synchronize()
    print(rank, f"4-byte memcpy->mc: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us each", flush=True)
# 4) unicast peer memcpy through symm_mem buffer_ptrs for comparison
peer = hdl.buffer_ptrs[(rank + 1) % world]
dist.barrier(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): cudart.cudaMemcpyAsync(peer + rank * chunk, src.data_ptr(), chunk, 4, st)
e1.record(); torch.cuda.synchronize()
print(rank, f"memcpy->peer 4 MiB: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us each", flush=True)
dist.barrier(); dist.destroy_process_group()
