import numpy as np, sys, os
from surely_raytracing_b200 import Scene, BuiltScene, capi
from oracle import orc
ok = True
for cfg in ["c5", "c1", "c2", "c3", "c4"]:
    b = BuiltScene(cfg, width=200, spp=16)
    g = Scene(b); o = orc.OracleScene(b, use_bvh=False)
    rays = g.camera_rays()
    for rep in range(3):
        h = g.trace(rays); hb = g.trace(rays, 1)
    ho = o.trace(rays)
    print(cfg, len(rays), "gpu-vs-oracle id mismatches", (h["prim"] != ho["prim"]).sum(), "bvh-vs-brute", (h["prim"] != hb["prim"]).sum(), flush=True)
    s, st = g.render(collect_stats=True)
    print(cfg, "render ok", s.mean() / 16, st, flush=True)
