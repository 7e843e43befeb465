"""Probe (2+ GPUs, torchrun): which route to peer-addressable device memory works on this box — torch symmetric memory, CUDA IPC."""
import os, sys, traceback
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
# --- route 1: torch.distributed._symmetric_memory
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    say("symm ok: buffer_ptrs", [hex(p) for p in h.buffer_ptrs][:4], "signal pads", len(h.signal_pad_ptrs), "multicast", getattr(h, "multicast_ptr", None))
    t.fill_(rank + 1)
    dist.barrier(); torch.cuda.synchronize()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.uint8)
    say("symm peer read", peer[:4].tolist())
except Exception as e:
    say("symm FAILED:", repr(e)); traceback.print_exc()
# --- route 2: CUDA IPC handles through torch storage sharing
try:
    x = torch.full((1 << 18,), float(rank + 1), device=dev)
    meta = x.untyped_storage()._share_cuda_()
    metas = [None] * world
    dist.all_gather_object(metas, meta)
    nxt = (rank + 1) % world
    st = torch.UntypedStorage._new_shared_cuda(*metas[nxt])
    y = torch.tensor([], dtype=torch.float32, device=torch.device("cuda", metas[nxt][0])).set_(st, 0, (16,))
    say("ipc peer tensor device", y.device, "values", y[:2].tolist())
    z = torch.zeros(16, device=dev); z.copy_(y); torch.cuda.synchronize()
    say("ipc copy ok", z[:2].tolist(), "ptr", hex(y.data_ptr()))
except Exception as e:
    say("ipc FAILED:", repr(e)); traceback.print_exc()
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
