"""torchrun probe: where does a multi-GPU step spend its time (search / all-gather / merge)?"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathway_b200.engine import BruteForceKnnMetricKind
from pathway_b200.external_index import B200BruteForceKNNIndex
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev); rank, world = dist.get_rank(), dist.get_world_size()
N, D, Q, K = 10_000_000 // world, 768, 1024, 10
ix = B200BruteForceKNNIndex(D, N, metric=BruteForceKnnMetricKind.COS, device=local)
for c in range(N // 250_000):
    g = torch.Generator(device=dev); g.manual_seed(c + 100 * rank)
    rows = torch.randn((250_000, D), generator=g, device=dev).to(torch.bfloat16)
    keys = np.zeros((250_000, 2), dtype=np.uint64); keys[:, 0] = np.arange(c * 250_000, (c + 1) * 250_000, dtype=np.uint64)
    torch.cuda.synchronize(dev); ix.add_device(keys, rows); del rows
q = torch.randn((Q, D), device=dev).to(torch.bfloat16)
packed = torch.empty((Q, K, 2), dtype=torch.int64, device=dev); gathered = torch.empty((world, Q, K, 2), dtype=torch.int64, device=dev)
oi = torch.empty((Q, K), dtype=torch.int64, device=dev); od = torch.empty((Q, K), dtype=torch.float64, device=dev)
ext = torch.cuda.ExternalStream(ix.stream_ptr, device=dev)
def run(mode, reps=30):
    with torch.cuda.stream(ext):
        for _ in range(5):
            ix.search_device_packed(q, K, 0, packed)
            if mode >= 1: dist.all_gather_into_tensor(gathered, packed)
            if mode >= 2: ix.merge_shards_device(gathered, oi, od)
        ext.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            ix.search_device_packed(q, K, 0, packed)
            if mode >= 1: dist.all_gather_into_tensor(gathered, packed)
            if mode >= 2: ix.merge_shards_device(gathered, oi, od)
        e1.record(ext); ext.synchronize()
    return e0.elapsed_time(e1) / reps
for mode, name in ((0, "search only"), (1, "+ all_gather"), (2, "+ merge"), (0, "search only again")):
    t = run(mode)
    if rank == 0: print(f"world {world}: {name}: {t:.3f} ms/step", flush=True)
ix.close(); dist.barrier(); dist.destroy_process_group()
