"""Which way of mapping a peer's buffer works in this environment?
  (a) torch.distributed._symmetric_memory (driver VMM handles)   (b) legacy CUDA IPC handles
torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import time
import traceback

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n = 32 << 20
peer = (rank + 1) % world

try:
    import torch.distributed._symmetric_memory as symm_mem
    buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    print(rank, "[symm] rendezvous ok", [hex(p) for p in hdl.buffer_ptrs], flush=True)
    hdl.barrier(channel=0)
    remote = hdl.get_buffer(peer, (n,), torch.float32)
    remote.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    print(rank, "[symm] peer write", bool((buf == float((rank - 1) % world + 1)).all()), flush=True)
except Exception:
    print(rank, "[symm] FAILED:\n" + traceback.format_exc()[-1500:], flush=True)

try:
    mine = torch.zeros(n, dtype=torch.float32, device=dev)
    st = mine.untyped_storage()
    h = st._share_cuda_()
    handles = [None] * world
    dist.all_gather_object(handles, h)
    ph = handles[peer]
    # (device, handle, storage_size_bytes, storage_offset_bytes, ref_counter_handle, ref_counter_offset, event_handle, event_sync_required)
    pst = torch.UntypedStorage._new_shared_cuda(dev.index, *ph[1:])
    remote = torch.empty(0, dtype=torch.float32, device=dev).set_(pst)
    src = torch.full((n,), float(rank + 1), device=dev)
    dist.barrier()
    for _ in range(3):
        remote.copy_(src)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        remote.copy_(src)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    dist.barrier()
    torch.cuda.synchronize()
    ok = bool((mine == float((rank - 1) % world + 1)).all())
    print(rank, "[ipc] peer write", ok, "P2P copy %.1f GB/s" % (n * 4 / ms / 1e6), "remote ptr", hex(remote.data_ptr()), flush=True)
except Exception:
    print(rank, "[ipc] FAILED:\n" + traceback.format_exc()[-1500:], flush=True)
dist.destroy_process_group()
