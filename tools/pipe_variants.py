"""Throughput of pipeline variants (which stage limits frames/s?)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N, pipeline, synth
W, H, D, BS = 1280, 720, 128, 9
K, Q = synth.camera_model(W, H)
maps = synth.warp_maps(W, H, 0) + synth.warp_maps(W, H, 1)
nfr = 32
base = [synth.stereo_pair(W, H, D, s) for s in range(4)]
L = np.stack([base[i % 4][0] for i in range(nfr)]); R = np.stack([base[i % 4][1] for i in range(nfr)])
ctx = N.Context(0)
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for name, kw in [("full", {}), ("no extractor", dict(extractor=-1)), ("no wls (1 matcher)", dict(use_wls=False)),
                 ("no wls, no extractor", dict(use_wls=False, extractor=-1)), ("mode 0 (5 paths) full", dict(mode=0))]:
    mode = kw.pop("mode", 1)
    cfg = pipeline.make_pipeline_config(W, H, D, BS, mode, Q, K, lanes=lanes, max_points=20000, **({"extractor": N.STEGER_IMPROVED} | kw))
    fp = pipeline.FramePipeline(cfg, maps=maps, ctx=ctx)
    dL, dR = fp.upload(L), fp.upload(R)
    for _ in range(2): fp.run_dev(dL, dR, nfr)
    ms = []
    for _ in range(3):
        fp.run_dev(dL, dR, nfr); ms.append(fp.last_ms)
    print("%-28s lanes %2d: %.3f ms/frame  (%.0f fps)" % (name, lanes, min(ms) / nfr, nfr / min(ms) * 1e3), flush=True)
    fp.close()
