import numpy as np, sys, time
from surely_raytracing_b200 import Scene, BuiltScene, capi
b = BuiltScene("c4")
g = Scene(b)
g.render(0, 100)
print("ROW6", file=sys.stderr, flush=True)
g.render(600, 700)
