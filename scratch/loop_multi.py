"""scratch: wall-clock anatomy of the sharded value loop (per rank): generate call vs the queries around it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mosaic-art-rust_b200")]
import numpy as np, torch, torch.distributed as dist
import mosaic_b200
from mosaic_b200 import synth
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
nid = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ids = [mosaic_b200.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0); nid = ids[0]
cfg = synth.CONFIGS["C3"]; gw, gh, n = cfg["grid_w"], cfg["grid_h"], cfg["n_materials"]
mats = torch.empty((n, 64, 64, 3), dtype=torch.uint8, pin_memory=True); mats.numpy()[...] = synth.materials(n)
tgt = torch.empty((cfg["target_h"], cfg["target_w"], 3), dtype=torch.uint8, pin_memory=True); tgt.numpy()[...] = synth.target(cfg["target_w"], cfg["target_h"])
ctx = mosaic_b200.Context(lr, rank, world, nid)
use_torch_stream = os.environ.get("OWN_STREAM", "0") != "1"
if use_torch_stream:
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
kw = dict(max_usage=3, adjacency_weight=0.3, color_strength=0.3, enable_optimization=True, optimization_iterations=1000, seed=42, build_similarity=True)
ctx.set_materials(mats); ctx.set_target(tgt, gw, gh)
for _ in range(3): ctx.generate(**kw)
if world > 1: dist.barrier()
torch.cuda.synchronize()
tg, tq = [], []
t00 = time.perf_counter()
for _ in range(20):
    t0 = time.perf_counter(); ctx.generate(**kw); t1 = time.perf_counter()
    for s in mosaic_b200.STAGES:
        try: ctx.stage_ms(s)
        except mosaic_b200.MosaicError: pass
    ctx.collective_ms(); t2 = time.perf_counter()
    tg.append(t1 - t0); tq.append(t2 - t1)
tot = (time.perf_counter() - t00) / 20
print(f"rank {rank}: per iteration {tot*1e3:.3f} ms; generate {np.median(tg)*1e3:.3f} (min {min(tg)*1e3:.3f} max {max(tg)*1e3:.3f}); queries {np.median(tq)*1e3:.3f}; torch_stream={use_torch_stream}", flush=True)
ctx.close()
if world > 1: dist.barrier(); dist.destroy_process_group()
