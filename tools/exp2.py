import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futspace_b200 as F
import tools.quick_time as Q
ctx = F.Context(0)
for flags, name in ((0, "tex"), (2, "tiled-ldg")):
    p = F.default_params(flags=flags)
    Q.run(ctx, 4096, 2160, 3840, 4000, prm=p, label="cfg3 " + name)
    Q.run(ctx, 2048, 1080, 1920, 2000, prm=p, label="cfg2 " + name)
    p = F.default_params(flags=flags, filter=0)
    Q.run(ctx, 4096, 2160, 3840, 4000, prm=p, label="cfg3-nearest " + name)
