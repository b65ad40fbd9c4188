import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers_modest import build_case, load_golden
from modest_b200 import pipeline as pl
name = sys.argv[1] if len(sys.argv) > 1 else "nusc_small"
case, shape = build_case(name); g = load_golden(name)
p = pl.SeedLabelPipeline(dict(plane_estimate=dict(range=[[-70, 70], [-20, 20]], max_hs=shape.max_hs, offset=0.05), image_shape=list(shape.image_shape)))
b = pl.make_batch([case.query], [g["pp"]], [case.calib])
labels_raw = torch.from_numpy(g["labels_raw"].astype(np.int32)).cuda()
n_clusters = torch.tensor([int(g["labels_raw"].max() + 1)], dtype=torch.int32, device="cuda")
plane2 = torch.from_numpy(g["plane2"][None].copy()).cuda()
lf, lfin, boxes, n_boxes, n_valid, flags = p.filter_and_fit(b, labels_raw, n_clusters, plane2)
nb = int(n_boxes.cpu()[0]); got = boxes.cpu().numpy()[0, :nb]
d = np.abs(got - g["boxes"])
np.set_printoptions(precision=6, suppress=True, linewidth=200)
for k in np.nonzero(d.max(axis=1) > 1e-9)[0]:
    print("box", k, "\n got", got[k], "\n ref", g["boxes"][k], "\n npts", (g["labels_final"] == k + 1).sum())
