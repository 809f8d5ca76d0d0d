import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futspace_b200 as F
import tools.quick_time as Q
ctx = F.Context(0)
Q.run(ctx, 4096, 2160, 3840, 4000, label="cfg3")
Q.run(ctx, 4096, 2160, 3840, 1, label="cfg3-dist1 (fixed cost)")
Q.run(ctx, 4096, 2160, 3840, 1000, label="cfg3-dist1000")
Q.run(ctx, 4096, 540, 3840, 4000, label="h540 (occupancy 32-40 warps)")
Q.run(ctx, 4096, 540, 3840, 1, label="h540-dist1")
Q.run(ctx, 4096, 1080, 3840, 4000, label="h1080")
p = F.default_params(filter=0)
Q.run(ctx, 4096, 2160, 3840, 4000, prm=p, label="cfg3-nearest")
Q.run(ctx, 4096, 540, 3840, 4000, prm=p, label="h540-nearest")
