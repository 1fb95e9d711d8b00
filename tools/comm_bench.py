"""Stand-alone bandwidth of the peer-memory all-reduce kernel (csrc/comm.cu) on an idle GPU: 48 MB chunks (the GradSync chunk size) and the
whole 1.44 GB 4M-B gradient arena, wide CTAs (512 threads) and slim CTAs (128 threads, option comm_slim) at several CTA counts.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/comm_bench.py
'bus GB/s' = bytes a rank moves over NVLink (reads + writes, 2 * (W-1)/W * size) per second."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
import torch
import torch.distributed as dist
from b200fm import lib
from b200fm.parallel import _P2PTransport

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 360 * (1 << 20)                                        # fp32 elements: 1.44 GB, the 4M-B arena
t = _P2PTransport(N, dev, None, 6)
t.arena.normal_()
stream = torch.cuda.current_stream()
seq = [0]


def run(n_elems, n_ctas, reps):
    chunks = max(1, N // n_elems)
    def once():
        for c in range(min(chunks, 8)):
            seq[0] += 1
            t.all_reduce(c * n_elems, n_elems, seq[0], stream, None, n_ctas)
    once()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        once()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * min(chunks, 8))
    tm = torch.tensor([ms], device=dev); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    return float(tm)


for slim in (0, 1):
    lib.set_option("comm_slim", slim)
    for n_ctas in ((2, 4, 6, 16, 64) if not slim else (24, 48, 64, 96, 128)):
        for mb in (48, 1440):
            n = mb * (1 << 20) // 4
            ms = run(n, n_ctas, 3 if mb == 48 else 2)
            bus = 2 * (world - 1) / world * n * 4 / (ms * 1e-3) / 1e9
            if rank == 0:
                print(f"{'slim' if slim else 'wide'} ctas={n_ctas:4d}  {mb:5d} MB: {ms * 1e3:9.1f} us  bus {bus:7.1f} GB/s  algo {n * 4 / (ms * 1e-3) / 1e9:7.1f} GB/s", flush=True)
# NCCL for comparison
x = torch.randn(48 * (1 << 20) // 4, device=dev)
for _ in range(3):
    dist.all_reduce(x)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
if rank == 0:
    ms = e0.elapsed_time(e1) / 10
    print(f"nccl all_reduce 48 MB: {ms * 1e3:.1f} us  algo {x.numel() * 4 / (ms * 1e-3) / 1e9:.1f} GB/s")
t.close()
dist.destroy_process_group()
