"""Where does the per-step cost of the output exchange come from?  (torchrun, N >= 2)

Runs the resident step of bench.py with pieces of the exchange switched off and prints the per-rank time of each
variant.  usage: python -m torch.distributed.run --nproc-per-node 2 ... tools/probe_gather.py [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from waternet_b200.api import Enhancer
from waternet_b200.dist import PeerGather
from waternet_b200.net import WaterNet

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
torch.cuda.set_device(local_rank)
device = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=device)
model = WaterNet()
model.load_state_dict(bench.bench_state_dict(), strict=True)
model = model.to(device).eval()
eng = Enhancer(model, device=device).engine
mode = model._mode()
B, H, W = 16, 1080, 1920
dev_in = torch.from_numpy(bench.synthetic_batch(B, H, W, seed=rank)).to(device)
dev_out = torch.empty_like(dev_in)
nb = eng.chunk_images(B, H, W)
gather = PeerGather.create((B, H, W, 3), torch.uint8, device)
assert isinstance(gather, PeerGather)
side = torch.cuda.Stream(device)
push_done = {}
copy_events = []
# a second, torch-managed buffer per rank exported the way torch.multiprocessing does it, for the comparison
from torch.multiprocessing.reductions import reduce_tensor
tbuf = torch.empty(world * gather.slot_bytes, dtype=torch.uint8, device=device)
handles = [None] * world
dist.all_gather_object(handles, reduce_tensor(tbuf))
peer_views = [tbuf if r == rank else fn(*a) for r, (fn, a) in enumerate(handles)]


def step(local=True, peer=True, finish=True, wait_push=True, events=True, torch_copy=False, once=False, pull=False, via_peer_ctx=False, fused=False, senders=None):
    cur = torch.cuda.current_stream(device)
    for a in range(0, B, nb):
        b = min(B, a + nb)
        if wait_push and a in push_done:
            cur.wait_event(push_done[a])
        if fused:  # the kernel that writes the pass's output stores it into every rank's block as well
            eng.enhance(dev_in[a:b], mode=mode, out_u8=dev_out[a:b], peer_out=gather.addresses(a))
            continue
        eng.enhance(dev_in[a:b], mode=mode, out_u8=dev_out[a:b])
        if not events:
            continue
        if once:  # one push of the whole batch after the last pass (not overlapped; one link wake-up per step)
            if b < B:
                continue
            a = 0
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            t0 = torch.cuda.Event(enable_timing=True)
            t0.record(side)
            off = rank * gather.slot_bytes + a * gather.row_bytes
            for r in range(world):
                p = (rank + r) % world
                if (p == rank and local) or (p != rank and peer and (senders is None or rank in senders)):
                    if pull and p != rank:  # this GPU's copy engine READS the peer's own slot (no ordering: cost probe only)
                        poff = p * gather.slot_bytes + a * gather.row_bytes
                        gather._check(gather._lib.wn_memcpy_async(gather._base + poff, gather._base_of(p) + poff,
                                                                  (b - a) * gather.row_bytes, side.cuda_stream), "memcpy")
                    elif via_peer_ctx and p != rank:  # same cudaMemcpyAsync, destination mapped in the context on the peer device
                        gather._check(gather._lib.wn_memcpy_async(peer_views[p].data_ptr() + off, dev_out[a:b].data_ptr(),
                                                                  (b - a) * gather.row_bytes, side.cuda_stream), "memcpy")
                    elif torch_copy:   # what a framework-level cross-device copy costs (orders itself on the peer GPU)
                        peer_views[p][off:off + (b - a) * gather.row_bytes].copy_(dev_out[a:b].view(-1), non_blocking=True)
                    else:
                        gather._check(gather._lib.wn_memcpy_async(gather._base_of(p) + off, dev_out[a:b].data_ptr(),
                                                                  (b - a) * gather.row_bytes, side.cuda_stream), "memcpy")
            done = torch.cuda.Event(enable_timing=True)
            done.record(side)
            push_done[a] = done
            copy_events.append((t0, done))
    if fused:
        gather.signal()
        with torch.cuda.stream(side):
            gather.wait()
    elif finish:
        with torch.cuda.stream(side):
            gather.finish()


import subprocess


def timed(kw):
    push_done.clear()
    copy_events.clear()
    dist.barrier()
    torch.cuda.synchronize()
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms",
                            "100", "-i", str(local_rank)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(**kw)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    smi.terminate()
    rows = [[float(x) for x in l.split(",")] for l in smi.communicate()[0].splitlines() if l.count(",") == 1]
    rows = [r for r in rows if r[1] > 500] or [[0.0, 0.0]]
    copy_ms = sum(a.elapsed_time(b) for a, b in copy_events) / max(1, len(copy_events))
    every = torch.zeros(world, 4, device=device)
    every[rank, 3] = copy_ms
    every[rank, 0] = e0.elapsed_time(e1) / steps
    every[rank, 1] = sum(r[0] for r in rows) / len(rows)
    every[rank, 2] = sum(r[1] for r in rows) / len(rows)
    dist.all_reduce(every)
    return [[round(float(v), 2) for v in row] for row in every.tolist()]


variants = [
    ("no exchange", dict(events=False, finish=False)),
    ("4 pushes, wn_memcpy_async", dict()),
    ("4 pushes, only rank 0 sends", dict(senders=(0,))),
    ("stores fused into the kernel", dict(fused=True)),
    ("4 pushes, tensor.copy_", dict(torch_copy=True)),
] * 4
for _ in range(3):
    step()
acc = {}
for name, kw in variants:
    ms = timed(kw)
    acc.setdefault(name, []).append([r[0] for r in ms])
    if rank == 0:
        print(f"{name:28s} per-rank [ms/step, mean SM MHz, mean W, ms per copy group] {ms}", flush=True)
if rank == 0:
    for name, v in acc.items():
        means = [round(sum(r[k] for r in v) / len(v), 2) for k in range(world)]
        print(f"SUMMARY {name:30s} per-rank ms/step, mean of {len(v)} rounds {means}   rounds {v}", flush=True)
gather.close()
dist.destroy_process_group()
