"""CPU-only: the travelling oracle against the golden vectors the unmodified reference produced
(tests/golden/*.npz, written by oracle/make_golden.py in the build container)."""
import numpy as np
import pytest

from oracle import modest_oracle as orc


def _cfg(shape):
    import json
    cfg = json.loads(json.dumps(orc.DEFAULT_MASK_CFG))
    cfg["plane_estimate"]["max_hs"] = shape.max_hs
    return cfg


@pytest.mark.parametrize("name", ["small", "nusc_small"])
def test_oracle_reproduces_reference_outputs(golden_case, name):
    case, shape, g = golden_case(name)
    counts = orc.neighbor_counts(case.query_fixed, case.history)
    assert np.array_equal(counts, g["counts"])
    sub = slice(0, 400)
    assert np.array_equal(orc.neighbor_counts_bruteforce(case.query_fixed[sub], case.history), g["counts"][sub])
    pp = orc.persistence_entropy(counts).astype(np.float32)
    assert np.array_equal(pp, g["pp"])
    cal = orc.Calib(table={"P2": g["calib_P2"], "Tr_velo_to_cam": g["calib_V2C"], "R0_rect": g["calib_R0"]})
    labels, objs, st = orc.seed_mask_for_scan(case.query, pp, cal, _cfg(shape), seed=1024 + case.scan_id,
                                              return_stages=True)
    assert np.array_equal(st["plane"], g["plane"]) and np.array_equal(st["plane2"], g["plane2"])
    assert np.array_equal(st["keep"], np.unpackbits(g["final_mask"])[:len(labels)].astype(bool))
    assert st["graph"].nnz == int(g["graph_nnz"])
    assert np.array_equal(st["raw"], g["labels_raw"])
    assert np.array_equal(labels, g["labels_final"])
    rows = np.array([[*o.t, o.l, o.w, o.h, o.ry, o.volume] for o in objs]).reshape(-1, 8)
    assert np.array_equal(rows, g["boxes"])
    gr = st["graph"].tocsr()
    assert np.array_equal(orc.dbscan_restated(gr.indptr, gr.indices, gr.data), g["labels_raw"][st["keep"]])
    iou = orc.bev_iou_matrix_f32(orc.boxes_for_nms(objs), orc.boxes_for_nms(objs))
    assert np.abs(iou - g["iou_cpu"]).max() <= 1e-4
    text, _ = orc.labels_for_scan(objs, cal, lambda b: g["iou_cpu"], image_shape=tuple(g["image_shape"]))
    assert text == str(g["label_text"])


def test_percentile_restatement_matches_numpy():
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 11, 257, 5000):
        v = rng.uniform(0, 1, n).astype(np.float32)
        for q in (0, 10, 20, 50, 73, 100):
            assert orc.percentile_f32_restated(v, q) == np.percentile(v, q)


def test_ransac_restatement_matches_sklearn():
    from sklearn.linear_model import RANSACRegressor
    rng = np.random.default_rng(1)
    for seed in range(6):
        n = 4000
        xy = rng.uniform(-30, 30, (n, 2)).astype(np.float32)
        z = (0.01 * xy[:, 0] - 0.02 * xy[:, 1] - 1.8 + rng.normal(0, 0.03, n)).astype(np.float32)
        z[rng.random(n) < 0.3] += rng.uniform(0.1, 2, 1)[0]
        np.random.seed(seed)
        model = RANSACRegressor().fit(xy, z)
        np.random.seed(seed)
        r = orc.ransac_restated(xy, z)
        assert r["n_trials"] == model.n_trials_
        assert (r["inlier_mask"] != model.inlier_mask_).sum() <= 2
        assert np.allclose(r["coef"], model.estimator_.coef_, atol=1e-6)


def test_ball_count_restatement_matches_ckdtree():
    """cKDTree's p=2 ball query == sequential f64 d2 <= r*r on f32-originated coordinates,
    including points planted a few ulps either side of the sphere surface."""
    from oracle import modest_oracle as orc
    rng = np.random.default_rng(11)
    q = rng.uniform(-3, 3, (400, 3)).astype(np.float32)
    hist = [rng.uniform(-3, 3, (m, 3)).astype(np.float32) for m in (0, 1, 700, 1500)]
    d = rng.normal(0, 1, (400, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    shell = np.float32(0.3) * (1 + rng.integers(-4, 5, (400, 1)) * 2.0 ** -22)
    hist.append((q.astype(np.float64) + d * shell).astype(np.float32))          # hits and misses by a hair
    assert np.array_equal(orc.neighbor_counts_bruteforce(q, hist), orc.neighbor_counts(q, hist))


def test_graph_and_dbscan_restatements_match_sklearn():
    """mutual-kNN AND radius edge set from first principles == the three sklearn graph calls the
    reference multiplies; union-find DBSCAN == sklearn.cluster.DBSCAN(metric='precomputed')."""
    from oracle import modest_oracle as orc
    rng = np.random.default_rng(12)
    blobs = [rng.normal(c, 0.4, (120, 3)) for c in ((0, 0, 0), (3, 0.5, 0), (0.5, 4, 0.3))]
    ptc = np.concatenate(blobs + [rng.uniform(-4, 8, (90, 3))]).astype(np.float32)
    pp = np.concatenate([rng.normal(m, 0.04, 120) for m in (0.2, 0.5, 0.8)] + [rng.uniform(0, 1, 90)]).astype(np.float32)
    for k, radius in ((70, 2.0), (10, 0.8)):
        G = orc.affinity_graph(ptc, pp, n_neighbors=k, radius=radius).tocsr()
        G.sort_indices()
        adj, w = orc.affinity_edges_bruteforce(ptc, pp, n_neighbors=k, radius=radius)
        rows = np.repeat(np.arange(G.shape[0]), np.diff(G.indptr))
        stored = np.zeros(adj.shape, dtype=bool)
        stored[rows, G.indices] = True                                # stored entries, whatever their weight
        assert np.array_equal(stored, adj)
        assert np.array_equal(G.data, w[rows, G.indices].astype(np.float64))
        for eps, ms in ((0.1, 10), (0.05, 4)):
            assert np.array_equal(orc.dbscan_restated(G.indptr, G.indices, G.data, eps, ms),
                                  orc.dbscan_labels(G, eps, ms))
