"""Probe (run under torchrun, 2+ GPUs): which way of getting peer-memory pointers works on this box?
(1) torch symmetric memory, (2) CUDA IPC handles of a cudaMalloc'd buffer."""
import ctypes, os, sys, traceback
import torch, torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD)
    say("symm ok: ptrs", [hex(p) for p in h.buffer_ptrs], "multicast", hex(getattr(h, "multicast_ptr", 0) or 0))
    t.zero_()
    h.barrier()
    peer = (rank + 1) % world
    h.get_buffer(peer, (1024,), torch.float32).fill_(float(rank + 1))
    h.barrier()
    say("symm peer store visible:", float(t[0]), "expected", float((rank - 1) % world + 1))
except Exception:
    say("symm FAILED"); traceback.print_exc()
try:
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(1 << 20)) == 0
    hbuf = (ctypes.c_ubyte * 64)()
    rc = rt.cudaIpcGetMemHandle(hbuf, p)
    say("cudaIpcGetMemHandle rc", rc)
    mine = torch.tensor(list(hbuf), dtype=torch.uint8, device=dev)
    allh = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allh, mine)
    peer = (rank + 1) % world
    ph = (ctypes.c_ubyte * 64)(*allh[peer].cpu().tolist())
    q = ctypes.c_void_p()
    rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_ubyte * 64, ctypes.c_uint]
    rc = rt.cudaIpcOpenMemHandle(ctypes.byref(q), ph, 1)
    say("cudaIpcOpenMemHandle rc", rc, hex(q.value or 0))
    src = torch.full((256,), float(rank + 1), device=dev)
    rt.cudaMemset(p, 0, ctypes.c_size_t(1024))
    torch.cuda.synchronize(); dist.barrier()
    rc = rt.cudaMemcpy(q, ctypes.c_void_p(src.data_ptr()), ctypes.c_size_t(1024), 3)
    torch.cuda.synchronize(); dist.barrier()
    out = torch.empty(256, device=dev)
    rt.cudaMemcpy(ctypes.c_void_p(out.data_ptr()), p, ctypes.c_size_t(1024), 3)
    say("ipc copy rc", rc, "value", float(out[0]), "expected", float((rank - 1) % world + 1))
    rt.cudaIpcCloseMemHandle(q)
    dist.barrier()
    rt.cudaFree(p)
except Exception:
    say("ipc FAILED"); traceback.print_exc()
dist.destroy_process_group()
