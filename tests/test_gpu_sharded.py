"""Multi-GPU parity (needs >= 2 B200s of one host; skipped otherwise): one process per GPU, sharded contexts
exchanging over NCCL, result identical to the single-GPU run."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "mosaic-art-rust_b200"))
import numpy as np, torch, torch.distributed as dist
import mosaic_b200
from mosaic_b200 import synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)   # only to hand the NCCL id around
ids = [mosaic_b200.Context.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
cfg = synth.CONFIGS["C2"]
gw, gh, n = cfg["grid_w"], cfg["grid_h"], cfg["n_materials"]
mats = synth.materials(n); tgt = synth.target(cfg["target_w"], cfg["target_h"])
kw = dict(max_usage=3, adjacency_weight=0.25, color_strength=0.4, enable_optimization=True,
          optimization_iterations=500, seed=7, build_similarity=True)
ctx = mosaic_b200.Context(rank, rank, world, ids[0])
ctx.set_materials(mats); ctx.set_target(tgt, gw, gh)
placed = np.zeros(gw * gh, np.int32); img = np.zeros(ctx.output_shape(), np.uint8)
rep = ctx.generate(placed_out=placed, image_out=img, **kw)
# the same job into a page-locked image: the ranks draw their rows of the GREEDY grid while rank 0 anneals and compose
# again only the cells the final grid changes (two-phase hand-off); then without annealing (one hand-off)
pin = mosaic_b200.PinnedArray(img.shape)
pin.array[...] = 0
placed2 = np.zeros(gw * gh, np.int32)
rep2 = ctx.generate(placed_out=placed2, image_out=pin.array, **kw)
ctx.sync()
img2 = pin.array.copy()
kw0 = dict(kw, enable_optimization=False, optimization_iterations=0)
pin.array[...] = 0
placed3 = np.zeros(gw * gh, np.int32)
ctx.generate(placed_out=placed3, image_out=pin.array, **kw0)
ctx.sync()
img3 = pin.array.copy()
placed4 = np.zeros(gw * gh, np.int32)
rep4 = ctx.generate(placed_out=placed4, **kw)      # image left on the device
sim = ctx.similarity_tri(); mlab = ctx.material_lab(); clab = ctx.cell_lab(); idx, d2 = ctx.candidates()
shard = ctx.shard()
ctx.close()
pin.close()
# single-GPU run of the same job on this rank's GPU
ref = mosaic_b200.Context(rank)
ref.set_materials(mats); ref.set_target(tgt, gw, gh)
rplaced = np.zeros(gw * gh, np.int32); rimg = np.zeros(ref.output_shape(), np.uint8)
rrep = ref.generate(placed_out=rplaced, image_out=rimg, **kw)
rsim = ref.similarity_tri(); rmlab = ref.material_lab(); rclab = ref.cell_lab(); ridx, rd2 = ref.candidates()
rplaced0 = np.zeros(gw * gh, np.int32); rimg0 = np.zeros(ref.output_shape(), np.uint8)
ref.generate(placed_out=rplaced0, image_out=rimg0, **kw0)
ref.close()
assert (mlab == rmlab).all() and (clab == rclab).all()
assert (idx == ridx).all() and (d2.view(np.uint32) == rd2.view(np.uint32)).all()
assert (sim.view(np.uint32) == rsim.view(np.uint32)).all()
assert (placed == rplaced).all() and rep.n_fallbacks == rrep.n_fallbacks and rep.opt.accepted_count == rrep.opt.accepted_count
rb, re_ = shard["rows"]; th = img.shape[0] // gh
assert (img[rb * th:re_ * th] == rimg[rb * th:re_ * th]).all()      # a sharded context delivers its own row band
assert (placed2 == rplaced).all() and rep2.opt.accepted_count == rrep.opt.accepted_count
assert (img2[rb * th:re_ * th] == rimg[rb * th:re_ * th]).all(), "two-phase (page-locked) rows differ"
assert (placed3 == rplaced0).all() and (img3[rb * th:re_ * th] == rimg0[rb * th:re_ * th]).all()
assert (placed4 == rplaced).all() and rep4.n_fallbacks == rrep.n_fallbacks
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_equals_single_gpu(lib_built, tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
        assert f"rank {r} ok" in o
