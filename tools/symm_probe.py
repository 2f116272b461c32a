"""Developer probe (2+ GPUs under torchrun): is torch symmetric memory usable for peer stores over NVLink?"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); ws = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty((ws, 1024), dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok: buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "signal_pad_ptrs", len(hdl.signal_pad_ptrs), flush=True)
    t.zero_()
    hdl.barrier()
    # write my row into every peer's buffer through their pointers
    for p in range(ws):
        peer = hdl.get_buffer(p, (ws, 1024), torch.float32)
        peer[rank].fill_(float(rank + 1))
    hdl.barrier()
    torch.cuda.synchronize()
    print(rank, "rows seen", t[:, 0].tolist(), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "symmetric memory unavailable:", repr(e), flush=True)
# fallback probe: plain CUDA IPC through torch.multiprocessing reductions is not needed if the above works
print(rank, "p2p access 0<->1:", torch.cuda.can_device_access_peer(0, 1) if ws > 1 else None, flush=True)
dist.barrier(); dist.destroy_process_group()
