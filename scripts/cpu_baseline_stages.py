"""Per-stage wall time of the reference's CPU path (oracle port: the same SciPy / scikit-learn /
NumPy calls the reference's programs make), BASELINE.md section 4 item 2 ("as-shipped mode": one
process, sklearn calls keep the reference's n_jobs=-1).  CPU only; prints a markdown table.

    python scripts/cpu_baseline_stages.py [n_scans]        (default 4; BASELINE.md plans 32)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modest_b200 import synth          # noqa: E402
from oracle import modest_oracle as orc  # noqa: E402
import bench                           # noqa: E402


def main():
    n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    acc = {}

    def timed(name, fn, *a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return out

    cfg = orc.DEFAULT_MASK_CFG
    pe, g, db, f = cfg["plane_estimate"], cfg["graph"], cfg["clustering"]["DBSCAN"], cfg["filtering"]
    t_all = time.perf_counter()
    for s in range(n_scans):
        case = synth.make_scan_case(200000 + s, synth.LYFT, n_traversals=bench.N_TRAV, n_points=bench.N_POINTS)
        t_scan = time.perf_counter()
        from scipy.spatial import cKDTree
        trees = timed("C  cKDTree build x16 (pre_compute_pp_score.py:188-190)", lambda: [cKDTree(h) for h in case.history])
        counts = timed("C  ball query x16 (:54-60)", lambda: np.stack(
            [t.query_ball_point(case.query_fixed[:, :3], r=0.3, return_length=True) for t in trees]).T)
        pp = timed("D  entropy (:68-75)", lambda: orc.persistence_entropy(counts).astype(np.float32))
        ptc = case.query
        cal = orc.Calib(table=case.calib)
        np.random.seed(1024 + case.scan_id)
        plane = timed("E  RANSAC plane (pointcloud_utils.py:44-65)", orc.fit_ground_plane, ptc[:, :3], max_hs=pe["max_hs"],
                      ptc_range=pe["range"])
        keep = timed("F,G masks (:68-81, generate_mask.py:57-65)", lambda: orc.keep_above_plane(
            ptc[:, :3], plane, offset=pe["offset"], only_range=pe["range"]) & orc.limit_range_mask(ptc, cfg["limit_range"]))
        graph = timed("H  kNN(70) x radius(2 m) graph + weights (clustering_utils.py:32-48)", orc.affinity_graph,
                      ptc[keep], pp[keep], g["n_neighbors"], g["radius"])
        raw = np.full(ptc.shape[0], -1, dtype=np.int64)
        raw[keep] = timed("I  DBSCAN precomputed (generate_mask.py:75-81)", orc.dbscan_labels, graph, db["eps"], db["min_samples"])
        labels, _ = timed("J  filter_labels incl. 2nd RANSAC (clustering_utils.py:94-135)", orc.filter_cluster_labels, ptc, pp,
                          raw, **f)
        rect = timed("K  velo -> rect (kitti_util.py:293-329)", cal.velo_to_rect, ptc[:, :3])
        objs = []
        for cid in range(1, labels.max() + 1):
            box = timed("L  closeness-to-edge box fit, 901 headings (pointcloud_utils.py:167-216,278-317)", orc.fit_box,
                        rect[labels == cid], rect)
            if f["min_volume"] < box.volume < f["max_volume"]:
                objs.append(box)
        timed("N,O BEV IoU + NMS + FOV + label text (pointcloud_utils.py:320-379)", orc.labels_for_scan, objs, cal,
              lambda b: orc.bev_iou_matrix_f32(b, b))
        acc["total"] = acc.get("total", 0.0) + time.perf_counter() - t_scan
    info = bench.host_info()
    print(f"Host: {info['cpu_model']}, {info['cpu_count']} logical CPUs; numpy {info['numpy']}, scipy {info['scipy']}, "
          f"scikit-learn {info['sklearn']}; {n_scans} synthetic Lyft-shape scans (60 000 points, 16 traversals x 1 frame), "
          f"one process, sklearn n_jobs=-1 where the reference sets it.\n")
    print("| Stage (reference file:line) | s / scan | share |")
    print("|---|---|---|")
    tot = acc.pop("total")
    for k, v in acc.items():
        print(f"| {k} | {v / n_scans:.3f} | {100 * v / tot:.1f} % |")
    print(f"| **total** | **{tot / n_scans:.2f}** ({n_scans / tot:.3f} scans/s) | |")
    print(f"\n(wall time of the whole run incl. synthetic data generation: {time.perf_counter() - t_all:.0f} s)")


if __name__ == "__main__":
    main()
