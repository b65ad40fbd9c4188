// TEST INFRASTRUCTURE ONLY.  Python binding for the reference's CPU IoU
// (generate_cluster_mask/utils/iou3d_nms/src/iou3d_cpu.cpp:232-252), compiled together with
// that file where it lies under /root/reference.  Needed because in the reference's full
// extension the host `inline` helpers of iou3d_cpu.cpp share their mangled names with the
// `__device__ inline` helpers of iou3d_nms_kernel.cu, and the linker may keep nvcc's host
// stubs of the latter (which call exit(1)) -- so boxes_iou_bev_cpu of the full build kills
// the process.  Building the CPU file alone avoids the clash without touching it.
#include <torch/extension.h>

int boxes_iou_bev_cpu(at::Tensor boxes_a_tensor, at::Tensor boxes_b_tensor, at::Tensor ans_iou_tensor);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("boxes_iou_bev_cpu", &boxes_iou_bev_cpu, "oriented boxes iou (reference CPU implementation)");
}
