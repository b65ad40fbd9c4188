"""CPU-only: host-side logic, the C ABI surface, and the numerical assumptions the kernels make
about numpy/OpenBLAS on the host that produced the goldens."""
import ctypes
import os
import re
import subprocess
import sys
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from modest_b200 import _lib
    header = open(os.path.join(ROOT, "include", "modest_b200.h")).read()
    declared = set(re.findall(r"\b(modest_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert handle.modest_abi_version() == 2


def test_missing_library_fails_loudly(monkeypatch):
    from modest_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmodest_b200.so")
    with pytest.raises(_lib.ModestError):
        _lib.lib()


def test_argument_errors_do_not_exit():
    from modest_b200 import _lib
    lib = _lib.lib()
    rc = lib.modest_plane_candidates_batch(None, 4, None, 1, -1.5, 0.0, 1.0, 0.0, 1.0, None, None, None, None)
    assert rc == _lib.ERR_ARG and b"null pointer" in lib.modest_last_error()
    num = ctypes.c_int(7)
    assert lib.modest_nms_bev(None, 0, 0.1, None, None, ctypes.byref(num), None, 0, None) == 0 and num.value == 0


def test_subset_draws_match_sklearn():
    from sklearn.utils.random import sample_without_replacement
    from modest_b200 import ransac_host
    for n in (5, 40, 250, 301, 20000):
        np.random.seed(n)
        want = [sample_without_replacement(n, 3, random_state=np.random.mtrand._rand) for _ in range(20)]
        np.random.seed(n)
        peek = ransac_host.peek_triples(n, 20)
        assert np.array_equal(np.array(want), peek)
        state = np.random.get_state()[1].copy()
        np.random.seed(n)
        assert np.array_equal(np.random.get_state()[1], state), "peek must not consume the stream"
        ransac_host.consume_trials(n, 7)
        np.random.seed(n)
        for _ in range(7):
            sample_without_replacement(n, 3, random_state=np.random.mtrand._rand)
        # both streams are now 7 draws in
        a = np.random.randint(1 << 30)
        np.random.seed(n)
        ransac_host.consume_trials(n, 7)
        assert np.random.randint(1 << 30) == a


def test_hydra_compat_composition(tmp_path):
    from modest_b200 import hydra_compat as h
    cfg_dir = os.path.join(ROOT, "modest_b200", "generate_cluster_mask", "configs")
    c = h.compose(cfg_dir, "generate_mask.yaml", ["data_root=/d", "plane_estimate.max_hs=-1.3",
                                                   "data_paths=nusc.yaml", "filtering.percentile=30"], cwd="/w")
    assert c.ptc_path == "/d/velodyne" and c.plane_estimate.max_hs == -1.3 and c.filtering.percentile == 30
    assert c.data_paths.idx_list == "/w/meta_data/nuscenes/train_idx.txt"
    assert c.data_paths.get("bbox_info_save_dst", "None").startswith("/w/")
    c = h.compose(cfg_dir, "generate_label_files.yaml", ["data_root=/d", "image_shape=[900, 1600]", "fov_only=False"])
    assert c.image_shape == [900, 1600] and c.fov_only is False and c.nms.threshold == 0.1
    with pytest.raises(h.MissingMandatoryValue):
        h.compose(cfg_dir, "pp_score.yaml", []).data_root
    text = h.OmegaConf.to_yaml(c)
    assert "image_shape" in text
    h.OmegaConf.save(c, str(tmp_path / "c.yaml"))
    assert (tmp_path / "c.yaml").read_text() == text


def test_reference_config_keys_are_all_present():
    """Every key of the reference's YAML files exists with the same default (drop-in CLI)."""
    import yaml
    ref_dir = "/root/reference/generate_cluster_mask/configs"
    if not os.path.isdir(ref_dir):
        pytest.skip("reference tree not present")
    ours = os.path.join(ROOT, "modest_b200", "generate_cluster_mask", "configs")
    for rel in ("pp_score.yaml", "generate_mask.yaml", "generate_label_files.yaml", "combine_labels.yaml", "data_paths/fw70_2m.yaml",
                "data_paths/nusc.yaml"):
        a = yaml.safe_load(open(os.path.join(ref_dir, rel)))
        b = yaml.safe_load(open(os.path.join(ours, rel)))
        assert a == b, rel


def test_blas_rounding_patterns_assumed_by_the_kernels():
    """csrc/boxes.cu and csrc/cluster.cu reproduce how numpy+OpenBLAS round small f64 products
    (k-ordered fused multiply-adds); check the host still behaves that way."""
    def fma(a, b, c):
        return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))
    rng = np.random.default_rng(0)
    A = rng.normal(size=(300, 2)) * 10
    th = 0.7
    rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    C = A @ rot.T
    ok = sum(fma(A[i, 1], rot.T[1, j], A[i, 0] * rot.T[0, j]) == C[i, j] for i in range(300) for j in range(2))
    p = (rng.normal(size=(300, 4)) * 20).astype(np.float32)
    pl = rng.normal(size=4)
    d = p[:, :3] @ pl[:3]
    ok2 = sum(fma(float(p[i, 2]), pl[2], fma(float(p[i, 0]), pl[0], float(p[i, 1]) * pl[1])) == d[i] for i in range(300))
    if ok != 600 or ok2 != 300:
        pytest.skip(f"host BLAS rounds differently ({ok}/600, {ok2}/300): footprint/plane knife-edge parity "
                    "with numpy is then not guaranteed on this host")


def test_synthetic_dataset_layout(tmp_path):
    import pickle
    from modest_b200 import synth
    info = synth.write_dataset(str(tmp_path / "d"), str(tmp_path / "m"), synth.LYFT, n_traversals=2,
                               frames_per_traversal=2, n_points=2000)
    assert len(info["idx"]) == 4
    pts = np.fromfile(tmp_path / "d" / "velodyne" / "000000.bin", dtype=np.float32).reshape(-1, 4)
    assert pts.shape == (2000, 4) and len(np.unique(pts[:, :3], axis=0)) == 2000
    vi = pickle.load(open(tmp_path / "m" / "valid_idx_info.pkl", "rb"))
    seq, frame, hist = vi[0]
    assert len(hist) == 2 and hist[0][0] == seq
    assert np.load(tmp_path / "d" / "l2e" / "000000.npy").dtype == np.float32
    assert open(tmp_path / "d" / "calib" / "000000.txt").read().startswith("P0:")


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as td
from modest_b200 import dist
dist.init(backend="gloo")
r, w = dist.rank(), dist.world_size()
ids = dist.shard(range(11))
local = {int(i): (b"scan %d from rank %d\n" % (i, r)) * (int(i) % 3) for i in ids}
allb = dist.gather_blobs(local)
assert sorted(allb) == list(range(11)), sorted(allb)
for i, v in allb.items():
    owner = [k for k in range(w) if i in np.array_split(np.arange(11), w)[k]][0]
    assert v == (b"scan %d from rank %d\n" % (i, owner)) * (i % 3)
assert dist.resolve_parts(1, 0) == (w, r) and dist.resolve_parts(4, 2) == (4, 2)
td.barrier()
print("ok", r)
'''


def test_two_rank_gloo_sharding_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_numa_binding_helper_is_a_noop_without_topology(tmp_path):
    """dist.bind_to_gpu_numa only narrows the CPU set when sysfs exposes one for the GPU; its
    cpulist parser follows the kernel's "a-b,c" grammar."""
    from modest_b200 import dist
    assert dist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert dist._parse_cpulist("") == set()
    assert dist.bind_to_gpu_numa(0, sysfs=str(tmp_path)) is None     # no GPU here / no such device directory
