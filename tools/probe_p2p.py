"""probe: do CUDA IPC peer mappings and torch symmetric memory work between two ranks on this box?"""
import os, sys, time, ctypes as C
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
class H(C.Structure): _fields_ = [("r", C.c_char * 64)]
p = C.c_void_p()
assert rt.cudaMalloc(C.byref(p), 1 << 26) == 0
h = H()
rc = rt.cudaIpcGetMemHandle(C.byref(h), p)
print(rank, "cudaIpcGetMemHandle rc", rc, flush=True)
mine = torch.frombuffer(bytearray(C.string_at(C.byref(h), 64)), dtype=torch.uint8).cuda()
allh = torch.empty(world * 64, dtype=torch.uint8, device="cuda")
dist.all_gather_into_tensor(allh, mine)
allh = allh.cpu().numpy().tobytes()
# fill my buffer with rank+1
rt.cudaMemset(p, rank + 1, 1 << 26); rt.cudaDeviceSynchronize()
dist.barrier()
peer = (rank + 1) % world
ph = H(); C.memmove(C.byref(ph), allh[peer * 64:(peer + 1) * 64], 64)
q = C.c_void_p()
rc = rt.cudaIpcOpenMemHandle(C.byref(q), ph, 1)
print(rank, "cudaIpcOpenMemHandle rc", rc, flush=True)
if rc == 0:
    dst = torch.empty(1 << 26, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), q, 1 << 26, 3, None)
    rt.cudaDeviceSynchronize()
    dt = time.perf_counter() - t0
    print(rank, "peer copy ok value", int(dst[0]), int(dst[-1]), "GB/s", 10 * (1 << 26) / dt / 1e9, flush=True)
can = C.c_int()
rt.cudaDeviceCanAccessPeer(C.byref(can), rank, peer); print(rank, "canAccessPeer", can.value, flush=True)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.float32, device=f"cuda:{rank}")
    hd = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "symm rendezvous ok", len(hd.buffer_ptrs), flush=True)
except Exception as e:
    print(rank, "symm failed:", repr(e)[:300], flush=True)
dist.barrier(); dist.destroy_process_group()
