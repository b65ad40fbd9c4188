"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Build-container only (needs /root/reference).  For every stage it (1) runs the reference's own
function (imported through oracle/reference_harness.py), (2) runs the restatement in
oracle/modest_oracle.py on the same input, (3) asserts they agree (bit-exact unless noted) and
(4) stores the reference's output.  The committed fixtures therefore pin the oracle to the
reference; tests/test_oracle_golden.py re-checks the oracle against them on any machine and
the `-m gpu` tests check the CUDA path against both.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from modest_b200 import synth  # noqa: E402
from oracle import modest_oracle as orc  # noqa: E402
from oracle import reference_harness as rh  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (scan_id, shape, n_points, n_traversals, frames_per_traversal)
    "small": (7, "lyft", 9000, 3, 1),
    "lyft60k_t2": (0, "lyft", 60000, 2, 1),        # BASELINE.json configs[0] shape
    "nusc_small": (11, "nusc", 8000, 3, 2),
}


def sha(a) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes() + str(a.dtype).encode() + str(a.shape).encode()).hexdigest()


def csr_digest(g) -> str:
    g = g.tocsr().copy()
    g.sort_indices()
    return sha(g.indptr.astype(np.int64)) + sha(g.indices.astype(np.int64)) + sha(g.data.astype(np.float64))


def build_case(name):
    scan_id, shape_name, n_pts, n_trav, fpt = CASES[name]
    shape = synth.NUSC if shape_name == "nusc" else synth.LYFT
    return synth.make_scan_case(scan_id, shape, n_traversals=n_trav, frames_per_traversal=fpt, n_points=n_pts), shape


def reference_outputs(ref, case, shape, cfg):
    """Bodies of the three reference CLI loops, each step calling the reference's own function."""
    from scipy.spatial import cKDTree
    from sklearn import cluster
    import warnings
    out = {}
    args = ref.AttrDict(max_neighbor_dist=0.3, ephe_type="entropy")
    trees = {t: cKDTree(h) for t, h in enumerate(case.history)}           # pre_compute_pp_score.py:188-190
    counts = ref.pp.count_neighbors(case.query_fixed, trees, args)         # :193
    H = ref.pp.compute_ephe_score(counts, args)                            # :194
    out["counts"], out["H"] = counts, H
    pp = H.astype(np.float32)
    ptc = case.query
    np.random.seed(synth.SEED_BASE + case.scan_id)
    pe = cfg["plane_estimate"]
    plane = ref.pc.estimate_plane(ptc[:, :3], max_hs=pe["max_hs"], ptc_range=pe["range"])   # generate_mask.py:55
    out["rng_after_plane1"] = np.random.get_state()[1][:4].copy(), np.random.get_state()[2]
    plane_mask = ref.pc.above_plane(ptc[:, :3], plane, offset=pe["offset"], only_range=pe["range"])
    lr = cfg["limit_range"]
    range_mask = (ptc[:, 0] <= lr[0][1]) * (ptc[:, 0] > lr[0][0]) * (ptc[:, 1] <= lr[1][1]) * (ptc[:, 1] > lr[1][0])
    final_mask = plane_mask * range_mask
    g = cfg["graph"]
    graph = ref.cl.precompute_affinity_matrix(ptc[final_mask], pp[final_mask], neighbor_type=g["neighbor_type"],
                                              affinity_type=g["affinity_type"], n_neighbors=g["n_neighbors"],
                                              radius=g["radius"])
    db = cfg["clustering"]["DBSCAN"]
    labels = np.zeros(ptc.shape[0], dtype=int) - 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        labels[final_mask] = cluster.DBSCAN(metric="precomputed", eps=db["eps"], min_samples=db["min_samples"],
                                            n_jobs=-1).fit(graph).labels_
    rng_state_before_filter = np.random.get_state()
    labels_filtered = ref.cl.filter_labels(ptc, pp, labels, **cfg["filtering"])
    np.random.set_state(rng_state_before_filter)
    plane2 = ref.pc.estimate_plane(ptc, max_hs=-1.5, ptc_range=((-70, 70), (-50, 50)))      # clustering_utils.py:126
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
        calib_path = fh.name
    synth.write_calib(calib_path, case.calib)
    calib = ref.ku.Calibration(calib_path)
    rect = calib.project_velo_to_rect(ptc[:, :3])
    objs = []
    lab2 = labels_filtered.copy()
    for i in range(1, lab2.max() + 1):
        obj = ref.pc.get_obj(rect[lab2 == i], rect, fit_method=cfg["bbox_gen"]["fit_method"])
        if obj.volume > cfg["filtering"]["min_volume"] and obj.volume < cfg["filtering"]["max_volume"]:
            objs.append(obj)
        else:
            lab2[lab2 == i] = 0
    mapping = {x: i for i, x in enumerate(sorted(list(set(lab2))))}
    for k in mapping:
        lab2[lab2 == k] = mapping[k]
    out.update(plane=plane, final_mask=final_mask, graph=graph, labels_raw=labels, labels_filtered=labels_filtered,
               plane2=plane2, labels_final=lab2, objs=objs, calib=calib, calib_path=calib_path, pp=pp)
    # gen_label_files.py:44-52 with the reference's CPU IoU op standing in for its CUDA op
    # (no GPU in the build container; the GPU tests compare against the real CUDA build)
    import torch
    from oracle import build_ref
    ext = build_ref.load_cpu()
    boxes = np.array([[o.t[0], o.t[2], 0, o.l, o.w, o.h, -o.ry] for o in objs])
    bt = torch.from_numpy(boxes).float()
    iou = torch.zeros((len(objs), len(objs)), dtype=torch.float32)
    if len(objs):
        ext.boxes_iou_bev_cpu(bt.contiguous(), bt.contiguous(), iou)
    iou = iou.numpy()
    keep = np.ones(len(objs), dtype=bool)
    order = np.diag(iou).argsort()[::-1] if len(objs) else []
    for idx in order:                                                       # pointcloud_utils.py:337-341
        if not keep[idx]:
            continue
        keep[iou[idx] > 0.1] = False
        keep[idx] = True
    kept_objs = [o for o, k in zip(objs, keep) if k]
    fov_objs = [o for o in kept_objs if ref.pc.is_within_fov(o, calib, list(shape.image_shape))]
    out.update(iou_cpu=iou, nms_keep=keep, label_text=ref.pc.objs2label(fov_objs, calib),
               fov_keep=np.array([ref.pc.is_within_fov(o, calib, list(shape.image_shape)) for o in objs]))
    return out


def objs_to_array(objs):
    return np.array([[*o.t, o.l, o.w, o.h, o.ry, o.volume] for o in objs], dtype=np.float64).reshape(-1, 8)


def main():
    ref = rh.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    import scipy
    import sklearn
    meta = dict(numpy=np.__version__, scipy=scipy.__version__, sklearn=sklearn.__version__, cases={})
    for name in CASES:
        case, shape = build_case(name)
        cfg = json.loads(json.dumps(orc.DEFAULT_MASK_CFG))
        cfg["plane_estimate"]["max_hs"] = shape.max_hs
        R = reference_outputs(ref, case, shape, cfg)
        N = case.query.shape[0]
        # ---------------- oracle vs reference, stage by stage ----------------
        o_counts = orc.neighbor_counts(case.query_fixed, case.history)
        assert np.array_equal(o_counts, R["counts"]), "counts"
        if N <= 10000:
            sub = slice(0, 1500)
            assert np.array_equal(orc.neighbor_counts_bruteforce(case.query_fixed[sub], case.history), R["counts"][sub])
        o_H = orc.persistence_entropy(o_counts)
        assert np.array_equal(o_H, R["H"]), "entropy"
        pp = R["pp"]
        o_labels, o_objs, st = orc.seed_mask_for_scan(case.query, pp, orc.Calib(R["calib_path"]), cfg,
                                                      seed=synth.SEED_BASE + case.scan_id, return_stages=True)
        assert np.array_equal(st["plane"], R["plane"]), "plane"
        assert np.array_equal(st["keep"], R["final_mask"]), "final_mask"
        assert csr_digest(st["graph"]) == csr_digest(R["graph"]), "graph"
        assert np.array_equal(st["raw"], R["labels_raw"]), "dbscan"
        assert np.array_equal(st["plane2"], R["plane2"]), "plane2"
        assert np.array_equal(o_labels, R["labels_final"]), "final labels"
        ra, oa = objs_to_array(R["objs"]), objs_to_array(o_objs)
        assert ra.shape == oa.shape and np.allclose(ra, oa, rtol=0, atol=1e-9), "boxes"
        exact_boxes = bool(np.array_equal(ra, oa))
        # restated third-party semantics
        g = st["graph"].tocsr()
        assert np.array_equal(orc.dbscan_restated(g.indptr, g.indices, g.data), R["labels_raw"][R["final_mask"]])
        if N <= 10000:
            kept = case.query[R["final_mask"]]
            adj, w = orc.affinity_edges_bruteforce(kept, pp[R["final_mask"]])
            dense = np.zeros(adj.shape, bool)
            rows = np.repeat(np.arange(g.shape[0]), np.diff(g.indptr))
            dense[rows, g.indices] = True
            assert np.array_equal(dense, adj), "edge set restatement"
            assert np.array_equal(g.data.astype(np.float32), w[rows, g.indices]), "weights restatement"
        for cid in range(min(5, R["labels_raw"].max() + 1)):
            v = pp[R["labels_raw"] == cid]
            assert orc.percentile_f32_restated(v, 20) == np.percentile(v, 20), "percentile restatement"
        iou_np = orc.bev_iou_matrix_f32(orc.boxes_for_nms(R["objs"]), orc.boxes_for_nms(R["objs"])) if len(R["objs"]) \
            else np.zeros((0, 0), np.float32)
        assert np.allclose(iou_np, R["iou_cpu"], atol=1e-4), "IoU restatement vs reference CPU op"  # thin boxes amplify f32 noise
        text, _ = orc.labels_for_scan(R["objs"], orc.Calib(R["calib_path"]), lambda b: R["iou_cpu"],
                                      image_shape=shape.image_shape)
        assert text == R["label_text"], "label text"
        # RANSAC trial-loop restatement against sklearn on the same stream
        np.random.seed(synth.SEED_BASE + case.scan_id)
        sel = case.query[orc.plane_candidates_mask(case.query[:, :3], cfg["plane_estimate"]["max_hs"],
                                                   cfg["plane_estimate"]["range"])]
        rr = orc.ransac_restated(sel[:, [0, 1]], sel[:, 2])
        np.random.seed(synth.SEED_BASE + case.scan_id)
        _, model = orc.fit_ground_plane(case.query[:, :3], cfg["plane_estimate"]["max_hs"],
                                        cfg["plane_estimate"]["range"], return_model=True)
        ransac_same_trials = bool(rr["n_trials"] == model.n_trials_)
        ransac_mask_diff = int((rr["inlier_mask"] != model.inlier_mask_).sum())
        # ---------------- f-1: combine_labels on synthetic detections ----------------
        drng = np.random.default_rng(4242 + case.scan_id)
        nd = len(R["objs"])
        loc = np.array([o.t for o in R["objs"]]).reshape(-1, 3) + drng.normal(0, 0.15, (nd, 3))
        dims = np.array([[o.l, o.h, o.w] for o in R["objs"]]).reshape(-1, 3) * drng.uniform(0.9, 1.15, (nd, 3))
        rys = np.array([o.ry for o in R["objs"]]) + drng.normal(0, 0.05, nd)
        n_fp = 12                                            # false positives on static structure / empty space
        rect_all = R["calib"].project_velo_to_rect(case.query[:, :3])
        fp_c = rect_all[drng.choice(len(rect_all), n_fp, replace=False)]
        loc = np.concatenate([loc, fp_c + np.array([0, 0.8, 0])])
        dims = np.concatenate([dims, np.column_stack([drng.uniform(3, 5, n_fp), drng.uniform(1.4, 2, n_fp), drng.uniform(1.5, 2.2, n_fp)])])
        rys = np.concatenate([rys, drng.uniform(-3, 3, n_fp)])
        preds = dict(frame_id="%06d" % case.scan_id, location=loc, dimensions=dims, rotation_y=rys,
                     score=drng.uniform(-0.5, 1.0, len(loc)))
        import copy
        ref_gate = np.array([ref.cb.filter_by_ppscore(rect_all, pp, o, percentile=50, threshold=0.5)
                             for o in ref.cb.predicts2objs(preds)])
        orc_gate = np.array([orc.detection_passes_pp_gate(rect_all, pp, o) for o in orc.detections_to_boxes(preds)])
        assert np.array_equal(ref_gate, orc_gate), "pp gate restatement"
        det_obj = [o for o, k in zip(ref.cb.predicts2objs(preds), ref_gate) if k & (o.score > -1)]
        gen_obj = copy.deepcopy(R["objs"])
        ref.cb.add_area_score(gen_obj)
        objs_all = det_obj + gen_obj
        import torch
        from oracle import build_ref
        ext = build_ref.load_cpu()
        bt = torch.from_numpy(np.array([[o.t[0], o.t[2], 0, o.l, o.w, o.h, -o.ry] for o in objs_all])).float()
        iou_m = torch.zeros((len(objs_all), len(objs_all)))
        ext.boxes_iou_bev_cpu(bt.contiguous(), bt.contiguous(), iou_m)
        iou_m = iou_m.numpy()
        keep_m = np.ones(len(objs_all), dtype=bool)
        for i_ in np.argsort([o.score for o in objs_all])[::-1]:        # objs_nms(use_score_rank=True)
            if keep_m[i_]:
                keep_m[iou_m[i_] > 0.1] = False
                keep_m[i_] = True
        merged = [o for o, k in zip(objs_all, keep_m) if k]
        merged = [o for o in merged if ref.pc.is_within_fov(o, R["calib"], list(shape.image_shape))]
        merge_text = ref.pc.objs2label(merged, R["calib"], with_score=True)
        o_text, _ = orc.merge_labels_for_scan(preds, copy.deepcopy(R["objs"]), rect_all, pp, orc.Calib(R["calib_path"]),
                                              lambda b: iou_m, image_shape=shape.image_shape, with_score=True)
        assert o_text == merge_text, "merge text restatement"
        # ---------------- f-3: road plane via the reference's own script ----------------
        import importlib.util
        import io
        import contextlib
        dp_dir = os.path.join(rh.REFERENCE_ROOT, "data_preprocessing")
        sys.path.insert(0, dp_dir)
        try:
            spec = importlib.util.spec_from_file_location("ref_road_ransac", os.path.join(dp_dir, "RANSAC.py"))
            ref_road = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ref_road)
        finally:
            sys.path.remove(dp_dir)
            sys.modules.pop("kitti_util", None)
        rd = tempfile.mkdtemp(prefix="modest_road_")
        os.makedirs(os.path.join(rd, "calib")); os.makedirs(os.path.join(rd, "velodyne"))
        synth.write_calib(os.path.join(rd, "calib", "000000.txt"), case.calib)
        case.query.tofile(os.path.join(rd, "velodyne", "000000.bin"))
        np.random.seed(77 + case.scan_id)
        with contextlib.redirect_stdout(io.StringIO()):
            ref_road.extract_ransac(os.path.join(rd, "calib"), os.path.join(rd, "velodyne"), os.path.join(rd, "planes"),
                                    min_h=shape.sensor_height - 0.3, max_h=shape.sensor_height + 0.3)
        road_text = open(os.path.join(rd, "planes", "000000.txt")).read()
        np.random.seed(77 + case.scan_id)
        ow, oh = orc.road_plane_for_scan(case.query, orc.Calib(R["calib_path"]), shape.sensor_height - 0.3,
                                         shape.sensor_height + 0.3)
        assert orc.road_plane_text(ow, oh) == road_text, "road plane restatement"
        import shutil
        shutil.rmtree(rd)
        g_sorted = g.copy()
        g_sorted.sort_indices()
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, f"{name}.npz"),
            scan_id=case.scan_id, n_points=N, query_sha=sha(case.query), history_sha=sha(np.concatenate(case.history)),
            counts=R["counts"].astype(np.int32), pp=pp, plane=R["plane"], plane2=R["plane2"],
            final_mask=np.packbits(R["final_mask"]), graph_digest=csr_digest(R["graph"]),
            graph_degree=np.diff(g_sorted.indptr).astype(np.int16), graph_nnz=g.nnz,
            labels_raw=R["labels_raw"].astype(np.int32), labels_filtered=R["labels_filtered"].astype(np.int32),
            labels_final=R["labels_final"].astype(np.int32), boxes=ra, iou_cpu=R["iou_cpu"], nms_keep=R["nms_keep"],
            fov_keep=R["fov_keep"], label_text=R["label_text"], n_trials1=model.n_trials_,
            inlier_count1=int(model.inlier_mask_.sum()), thr1=np.float32(orc.mad_threshold(sel[:, 2])),
            calib_P2=np.asarray(case.calib["P2"]), calib_V2C=np.asarray(case.calib["Tr_velo_to_cam"]),
            calib_R0=np.asarray(case.calib["R0_rect"]), image_shape=np.array(shape.image_shape),
            max_hs=shape.max_hs, det_location=preds["location"], det_dimensions=preds["dimensions"],
            det_rotation_y=preds["rotation_y"], det_score=preds["score"], det_pp_gate=ref_gate,
            merge_iou_cpu=iou_m, merge_text=merge_text, road_text=road_text, road_plane=np.array([*ow, oh]))
        meta["cases"][name] = dict(n_points=N, n_kept=int(R["final_mask"].sum()), graph_nnz=int(g.nnz),
                                   n_clusters_raw=int(R["labels_raw"].max() + 1), n_boxes=len(R["objs"]),
                                   n_labels=R["label_text"].count("\n") + (1 if R["label_text"] else 0),
                                   oracle_boxes_bit_exact=exact_boxes, ransac_restated_same_trials=ransac_same_trials,
                                   ransac_restated_mask_diff=ransac_mask_diff)
        print(name, meta["cases"][name])
        os.unlink(R["calib_path"])
    with open(os.path.join(GOLDEN_DIR, "MANIFEST.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


if __name__ == "__main__":
    main()
