"""TEST INFRASTRUCTURE ONLY -- build the reference's own iou3d_nms_cuda extension, unmodified.

Compiles the four sources where they lie under
/root/reference/generate_cluster_mask/utils/iou3d_nms/src/ (iou3d_cpu.cpp, iou3d_nms_api.cpp,
iou3d_nms.cpp, iou3d_nms_kernel.cu -- the list in that directory's setup.py:6-11) for sm_100
with torch.utils.cpp_extension, writing only into oracle/_ref/.  No reference source is copied
into this repository; oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.

The resulting module is (a) the like-for-like GPU oracle for the BEV IoU stage in the `-m gpu`
tests and (b) the CPU IoU (`boxes_iou_bev_cpu`) used by oracle/make_golden.py.

Run:  python oracle/build_ref.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/generate_cluster_mask/utils/iou3d_nms/src"
NAMES = ["iou3d_cpu.cpp", "iou3d_nms_api.cpp", "iou3d_nms.cpp", "iou3d_nms_kernel.cu"]


def built_path():
    p = os.path.join(OUT, "iou3d_nms_cuda.so")
    return p if os.path.exists(p) else None


def build(verbose=False):
    if built_path():
        return built_path()
    if not os.path.isdir(SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    cpp_extension.load(
        name="iou3d_nms_cuda", sources=[os.path.join(SRC, n) for n in NAMES],
        extra_cflags=["-g"], extra_cuda_cflags=["-O2"],   # the flags of the reference's setup.py
        build_directory=OUT, verbose=verbose, is_python_module=False)
    return built_path()


def _import_so(name, p):
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Import the built CUDA module (None when it was never built).  Use its *_gpu functions
    only: its boxes_iou_bev_cpu exits the process (see ref_cpu_binding.cpp)."""
    p = built_path()
    return None if p is None else _import_so("iou3d_nms_cuda", p)


def cpu_built_path():
    p = os.path.join(OUT, "cpu", "iou3d_ref_cpu.so")
    return p if os.path.exists(p) else None


def build_cpu(verbose=False):
    """The reference's iou3d_cpu.cpp alone + oracle/ref_cpu_binding.cpp -> oracle/_ref/cpu/."""
    if cpu_built_path():
        return cpu_built_path()
    if not os.path.isdir(SRC):
        return None
    out = os.path.join(OUT, "cpu")
    os.makedirs(out, exist_ok=True)
    from torch.utils import cpp_extension
    cpp_extension.load(name="iou3d_ref_cpu",
                       sources=[os.path.join(SRC, "iou3d_cpu.cpp"), os.path.join(HERE, "ref_cpu_binding.cpp")],
                       extra_cflags=["-g"], extra_include_paths=["/usr/local/cuda/include"],
                       build_directory=out, verbose=verbose, is_python_module=False)
    return cpu_built_path()


def load_cpu():
    p = cpu_built_path()
    return None if p is None else _import_so("iou3d_ref_cpu", p)


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
    print(build_cpu(verbose="-v" in sys.argv))
