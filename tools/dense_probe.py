"""Diagnostic: the fold kernels in the dense regime of a sharded run -- rank 0's shard of the 8-GPU weak-scaling
workload (512 columns x 32768 rows, all particles inside one eighth of the grid), folded locally on one GPU."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import tendrils_b200 as T
from tendrils_b200.spawn import PixelSpawner
from tendrils_b200.spawn.pixels import pixelsFrag, mat3_identity, mat3_scale
from util import synthetic_image

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
R, G = 4096, 1024
t = T.Tendrils(T.Device(G, G, rank=0, world_size=world))
t.setup([R, R * world]); t.resize()
t.gl.world_size = 1                       # fold locally: no ring
sp = PixelSpawner(t.gl, {"shader": pixelsFrag, "buffer": synthetic_image(G, G), "speed": 0.3, "jitterRad": 2})
sp.spawnMatrix = mat3_scale(mat3_identity(), [-1, 1])
sp.spawn(t)
for k in range(steps):
    t.timer.tick(); t.step().draw()
t.particles.sync(); t.particles.timing(reset=True)
for k in range(10):
    t.timer.tick(); t.step().draw()
tm = t.particles.timing()
print("dense probe world=%d: integrate %.0f us, splat %.0f us, fragments %d" % (
    world, 1e3 * tm["integrate_ms"] / tm["n_integrate"], 1e3 * tm["splat_ms"] / tm["n_splat"], t.particles.stats()["last_fragments"]))
