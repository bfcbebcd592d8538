"""Run the instrumented CPU oracle on a BASELINE.json config at its exact resolution and store the
reference-control-flow counters (tree nodes, shadow rays, intersectSphere evaluations T) that define
the Mrays/s numerator and the roofline's algorithmic flops 15*T.  The synthetic scene
is seeded, so these are constants of the workload; bench.py reads them from tests/golden/."""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rt = importlib.import_module("html5-canvas-raytracer_b200")
from oracle import oracle  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3_synth1024_4k"
scene, W, H, segs, g = rt.scenes.config_workload(name)
objs = rt.scenes.demo_objects() if scene == "demo" else rt.scenes.synth_objects(1024, 1024)
sd = rt.SceneDesc(objs)
cam = rt.scenes.demo_camera(W, H)
t = time.time()
img, cnt = oracle.render(sd, cam, rt.make_opts(segs=segs, spp_grid=g))
dt = time.time() - t
cnt.update({"config": name, "width": W, "height": H, "segs": segs, "spp_grid": g, "n_spheres": len(objs),
            "oracle_seconds": round(dt, 2), "oracle_threads": oracle.max_threads()})
path = os.path.join(ROOT, "tests", "golden", f"counters_{name}.json")
json.dump(cnt, open(path, "w"), indent=1)
import numpy as np
np.savez_compressed(f"/tmp/work/oracle_{name}.npz", img=img)
print(json.dumps(cnt))
