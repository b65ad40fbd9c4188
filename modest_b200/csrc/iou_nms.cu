// Stage N: rotated bird's-eye-view IoU, overlap and NMS for boxes [x, y, z, dx, dy, dz, heading].
//
// Replaces the reference's only native op, module `iou3d_nms_cuda`
// (utils/iou3d_nms/src/iou3d_nms_api.cpp:11-17):
//   boxes_overlap_bev_gpu / boxes_iou_bev_gpu   iou3d_nms.cpp:48-88  -> kernels at
//                                               iou3d_nms_kernel.cu:236-265
//   nms_gpu / nms_normal_gpu                    iou3d_nms.cpp:90-188 -> kernels :267-372
// and the greedy loop of objs_nms() (utils/pointcloud_utils.py:320-344).
//
// The polygon-clipping arithmetic has to reproduce the reference kernel's float32 results
// bit for bit (objs_nms orders boxes by the rounding noise of their self-IoU),
// so the geometric recipe is the same -- rotate the 4 corners, collect edge/edge intersections
// and corners inside the other box (1e-2 margin), order around the centroid by atan2 with a
// bubble sort, shoelace -- and every expression keeps the reference's operand order.  What is
// new: one warp-friendly kernel for all pair shapes, a device-side greedy reduction (no
// blocking D2H of the suppression mask unless the legacy entry point asks for host output),
// and a batched "seed NMS" that handles every scan of a batch in one launch.
#include "common.cuh"

namespace modest {
extern void note_launch(int n);

namespace bev {

constexpr float kEps = 1e-8f;
constexpr float kMargin = 1e-2f;

struct P2 { float x, y; };

__device__ __forceinline__ float cross3(const P2& p1, const P2& p2, const P2& p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
__device__ __forceinline__ float cross2(const P2& a, const P2& b) { return a.x * b.y - a.y * b.x; }

__device__ __forceinline__ bool bbox_overlap(const P2& p1, const P2& p2, const P2& q1, const P2& q2) {
  return min(p1.x, p2.x) <= max(q1.x, q2.x) && min(q1.x, q2.x) <= max(p1.x, p2.x) &&
         min(p1.y, p2.y) <= max(q1.y, q2.y) && min(q1.y, q2.y) <= max(p1.y, p2.y);
}

__device__ __forceinline__ bool inside_box(const float* box, const P2& p) {
  const float cx = box[0], cy = box[1];
  const float ac = cos(-box[6]), as = sin(-box[6]);
  const float rx = (p.x - cx) * ac + (p.y - cy) * (-as);
  const float ry = (p.x - cx) * as + (p.y - cy) * ac;
  return fabs(rx) < box[3] / 2 + kMargin && fabs(ry) < box[4] / 2 + kMargin;
}

__device__ __forceinline__ bool segment_hit(const P2& p1, const P2& p0, const P2& q1, const P2& q0, P2* out) {
  if (!bbox_overlap(p0, p1, q0, q1)) return false;
  const float s1 = cross3(q0, p1, p0);
  const float s2 = cross3(p1, q1, p0);
  const float s3 = cross3(p0, q1, q0);
  const float s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
  const float s5 = cross3(q1, p1, p0);
  if (fabs(s5 - s1) > kEps) {
    out->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    out->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float D = a0 * b1 - a1 * b0;
    out->x = (b0 * c1 - b1 * c0) / D;
    out->y = (a1 * c0 - a0 * c1) / D;
  }
  return true;
}

__device__ __forceinline__ void turn(const P2& c, float ac, float as, P2* p) {
  const float nx = (p->x - c.x) * ac + (p->y - c.y) * (-as) + c.x;
  const float ny = (p->x - c.x) * as + (p->y - c.y) * ac + c.y;
  p->x = nx; p->y = ny;
}

__device__ __forceinline__ void corners_of(const float* b, P2* c /*[5]*/) {
  const float hx = b[3] / 2, hy = b[4] / 2;
  const float x1 = b[0] - hx, y1 = b[1] - hy, x2 = b[0] + hx, y2 = b[1] + hy;
  c[0] = P2{x1, y1}; c[1] = P2{x2, y1}; c[2] = P2{x2, y2}; c[3] = P2{x1, y2};
}

// Boxes whose circumscribed circles are clearly apart share no point: overlap_area would build an
// empty intersection polygon and return +0 (5 cm of slack covers the 1e-2 margin of inside_box and
// any float32 rounding; NaNs compare false and take the full path).  Kept OUT of overlap_area and
// that function out of line, so that its float32 code -- which has to reproduce the reference
// kernel's bits -- does not depend on what surrounds the call.
__device__ __forceinline__ bool clearly_apart(const float* box_a, const float* box_b) {
  const float dx = box_a[0] - box_b[0], dy = box_a[1] - box_b[1];
  const float ra = 0.5f * sqrtf(box_a[3] * box_a[3] + box_a[4] * box_a[4]);
  const float rb = 0.5f * sqrtf(box_b[3] * box_b[3] + box_b[4] * box_b[4]);
  const float rs = ra + rb + 0.05f;
  return dx * dx + dy * dy > 1.001f * rs * rs;
}

__device__ __noinline__ float overlap_area_full(const float* box_a, const float* box_b) {
  const float a_angle = box_a[6], b_angle = box_b[6];
  P2 ca[5], cb[5];
  corners_of(box_a, ca);
  corners_of(box_b, cb);
  const P2 centre_a{box_a[0], box_a[1]}, centre_b{box_b[0], box_b[1]};
  const float a_cos = cos(a_angle), a_sin = sin(a_angle);
  const float b_cos = cos(b_angle), b_sin = sin(b_angle);
  for (int k = 0; k < 4; ++k) {
    turn(centre_a, a_cos, a_sin, &ca[k]);
    turn(centre_b, b_cos, b_sin, &cb[k]);
  }
  ca[4] = ca[0];
  cb[4] = cb[0];

  P2 poly[16];
  P2 centre{0.f, 0.f};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (segment_hit(ca[i + 1], ca[i], cb[j + 1], cb[j], &poly[cnt])) {
        centre.x = centre.x + poly[cnt].x;
        centre.y = centre.y + poly[cnt].y;
        ++cnt;
      }
  for (int k = 0; k < 4; ++k) {
    if (inside_box(box_a, cb[k])) {
      centre.x = centre.x + cb[k].x; centre.y = centre.y + cb[k].y;
      poly[cnt++] = cb[k];
    }
    if (inside_box(box_b, ca[k])) {
      centre.x = centre.x + ca[k].x; centre.y = centre.y + ca[k].y;
      poly[cnt++] = ca[k];
    }
  }
  centre.x /= cnt;
  centre.y /= cnt;

  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (atan2(poly[i].y - centre.y, poly[i].x - centre.x) > atan2(poly[i + 1].y - centre.y, poly[i + 1].x - centre.x)) {
        const P2 t = poly[i]; poly[i] = poly[i + 1]; poly[i + 1] = t;
      }

  float area = 0;
  for (int k = 0; k < cnt - 1; ++k) {
    const P2 u{poly[k].x - poly[0].x, poly[k].y - poly[0].y};
    const P2 v{poly[k + 1].x - poly[0].x, poly[k + 1].y - poly[0].y};
    area += cross2(u, v);
  }
  return fabs(area) / 2.0;
}

__device__ __forceinline__ float overlap_area(const float* a, const float* b) {
  return clearly_apart(a, b) ? 0.f : overlap_area_full(a, b);
}

__device__ __forceinline__ float iou(const float* a, const float* b) {
  const float sa = a[3] * a[4];
  const float sb = b[3] * b[4];
  const float ov = overlap_area(a, b);
  return ov / fmaxf(sa + sb - ov, kEps);
}

__device__ __forceinline__ float iou_axis_aligned(const float* a, const float* b) {
  const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
  const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
  const float wd = fmaxf(right - left, 0.f), ht = fmaxf(bottom - top, 0.f);
  const float inter = wd * ht;
  const float sa = a[3] * a[4], sb = b[3] * b[4];
  return inter / fmaxf(sa + sb - inter, kEps);
}

}  // namespace bev

// mode 0: IoU, mode 1: raw overlap area
__global__ void __launch_bounds__(128) bev_pairs_kernel(int na, const float* __restrict__ a, int nb,
                                                        const float* __restrict__ b, float* __restrict__ out, int mode) {
  const long long total = (long long)na * nb;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / nb), j = (int)(e % nb);
    out[e] = mode == 0 ? bev::iou(a + 7 * i, b + 7 * j) : bev::overlap_area(a + 7 * i, b + 7 * j);
  }
}

// suppression bit matrix: bit j of mask[i][j/64] = (j > i) && IoU(i,j) > thr
__global__ void __launch_bounds__(64) nms_mask_kernel(int n, float thr, const float* __restrict__ boxes,
                                                      unsigned long long* __restrict__ mask, int rotated) {
  const int row_blk = blockIdx.y, col_blk = blockIdx.x;
  const int rows = min(n - row_blk * 64, 64), cols = min(n - col_blk * 64, 64);
  __shared__ float tile[64 * 7];
  for (int k = threadIdx.x; k < cols * 7; k += 64) tile[k] = boxes[(size_t)col_blk * 64 * 7 + k];
  __syncthreads();
  if ((int)threadIdx.x < rows) {
    const int i = row_blk * 64 + threadIdx.x;
    const float* me = boxes + (size_t)i * 7;
    unsigned long long bits = 0;
    const int start = row_blk == col_blk ? threadIdx.x + 1 : 0;
    for (int j = start; j < cols; ++j) {
      const float v = rotated ? bev::iou(me, tile + j * 7) : bev::iou_axis_aligned(me, tile + j * 7);
      if (v > thr) bits |= 1ull << j;
    }
    mask[(size_t)i * ((n + 63) / 64) + col_blk] = bits;
  }
}

// greedy sweep over the bit matrix in index order (boxes arrive sorted by score), one warp
__global__ void nms_reduce_kernel(int n, const unsigned long long* __restrict__ mask, long long* __restrict__ keep,
                                  int* __restrict__ num_keep, unsigned long long* __restrict__ removed) {
  const int blocks = (n + 63) / 64;
  const int lane = threadIdx.x;
  for (int b = lane; b < blocks; b += 32) removed[b] = 0;
  __syncwarp();
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    const bool alive = !((removed[i / 64] >> (i % 64)) & 1ull);
    if (alive) {
      if (lane == 0) keep[kept] = i;
      ++kept;
      for (int b = i / 64 + lane; b < blocks; b += 32) removed[b] |= mask[(size_t)i * blocks + b];
    }
    __syncwarp();
  }
  if (lane == 0) *num_keep = kept;
}

// ---- batched seed NMS (objs_nms with use_score_rank=False), one CTA per scan --------------------
// boxes64: (S, max_boxes, 8) f64 rows [t.x, t.y, t.z, l, w, h, ry, volume] from the box-fit stage.
__global__ void __launch_bounds__(256) seed_nms_kernel(const double* __restrict__ boxes64, const int32_t* __restrict__ n_boxes,
                                                       int max_boxes, float thr, float* __restrict__ iou_out /* (S,max,max) or NULL */,
                                                       float* __restrict__ iou_ws /* (S,max,max) */, uint8_t* __restrict__ keep_out) {
  const int s = blockIdx.x;
  const int K = n_boxes[s];
  extern __shared__ float sh[];
  float* bx = sh;                          // K*7
  int* order = reinterpret_cast<int*>(sh + 7 * max_boxes);
  float* m = iou_ws + (size_t)s * max_boxes * max_boxes;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const double* b = boxes64 + ((size_t)s * max_boxes + k) * 8;
    // [t.x, t.z, 0, l, w, h, -ry] rounded to float32 (pointcloud_utils.py:322-324)
    bx[7 * k + 0] = (float)b[0]; bx[7 * k + 1] = (float)b[2]; bx[7 * k + 2] = 0.f;
    bx[7 * k + 3] = (float)b[3]; bx[7 * k + 4] = (float)b[4]; bx[7 * k + 5] = (float)b[5];
    bx[7 * k + 6] = (float)(-b[6]);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K * K; e += blockDim.x) {
    const int i = e / K, j = e % K;
    const float v = bev::iou(bx + 7 * i, bx + 7 * j);
    m[i * max_boxes + j] = v;
    if (iou_out) iou_out[(size_t)s * max_boxes * max_boxes + i * max_boxes + j] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // order = argsort(diag)[::-1] with a stable ascending sort (numpy's insertion sort for
    // K <= 16): descending self-IoU, ties in descending index order.
    for (int k = 0; k < K; ++k) order[k] = k;
    for (int a = 1; a < K; ++a) {
      const int v = order[a];
      const float dv = m[v * max_boxes + v];
      int p = a - 1;
      while (p >= 0 && m[order[p] * max_boxes + order[p]] > dv) { order[p + 1] = order[p]; --p; }
      order[p + 1] = v;
    }
    uint8_t* keep = keep_out + (size_t)s * max_boxes;
    for (int k = 0; k < max_boxes; ++k) keep[k] = k < K ? 1 : 0;
    for (int r = K - 1; r >= 0; --r) {
      const int i = order[r];
      if (!keep[i]) continue;
      for (int j = 0; j < K; ++j) if (m[i * max_boxes + j] > thr) keep[j] = 0;
      keep[i] = 1;
    }
  }
}

}  // namespace modest

using namespace modest;

static int pairs_launch(const float* a, int na, const float* b, int nb, float* out, int mode, cudaStream_t stream) {
  if (na <= 0 || nb <= 0) return MODEST_OK;
  MODEST_REQUIRE(a && b && out, "boxes_bev: null pointer argument");
  long long total = (long long)na * nb;
  long long blocks = (total + 127) / 128;
  if (blocks > 148 * 64) blocks = 148 * 64;
  bev_pairs_kernel<<<(unsigned)blocks, 128, 0, stream>>>(na, a, nb, b, out, mode);
  MODEST_LAUNCH_CHECK("bev_pairs_kernel");
  note_launch(1);
  return MODEST_OK;
}

extern "C" int modest_boxes_iou_bev(const float* d_boxes_a, int num_a, const float* d_boxes_b, int num_b,
                                    float* d_iou, void* stream) {
  return pairs_launch(d_boxes_a, num_a, d_boxes_b, num_b, d_iou, 0, static_cast<cudaStream_t>(stream));
}
extern "C" int modest_boxes_overlap_bev(const float* d_boxes_a, int num_a, const float* d_boxes_b, int num_b,
                                        float* d_overlap, void* stream) {
  return pairs_launch(d_boxes_a, num_a, d_boxes_b, num_b, d_overlap, 1, static_cast<cudaStream_t>(stream));
}

extern "C" size_t modest_nms_workspace_bytes(int n) {
  const size_t blocks = (size_t)(n + 63) / 64;
  return align_up(sizeof(unsigned long long) * (size_t)n * blocks, 256) + align_up(sizeof(unsigned long long) * blocks, 256) +
         align_up(sizeof(long long) * (size_t)n, 256) + 512;
}

static int nms_impl(const float* d_boxes, int n, float thr, long long* d_keep, long long* h_keep, int* h_num,
                    void* d_ws, size_t ws_bytes, cudaStream_t stream, int rotated) {
  MODEST_REQUIRE(h_num, "nms: h_num_out is NULL");
  *h_num = 0;
  if (n <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_boxes && d_ws, "nms: null pointer argument");
  MODEST_REQUIRE(ws_bytes >= modest_nms_workspace_bytes(n), "nms: workspace too small");
  const int blocks = (n + 63) / 64;
  Arena ar(d_ws, ws_bytes);
  unsigned long long* mask = ar.take<unsigned long long>((size_t)n * blocks);
  unsigned long long* removed = ar.take<unsigned long long>(blocks);
  long long* keep_ws = ar.take<long long>(n);
  int* num_dev = reinterpret_cast<int*>(ar.take<long long>(1));
  MODEST_REQUIRE(ar.ok(), "workspace too small for the requested sizes");
  long long* keep = d_keep ? d_keep : keep_ws;
  nms_mask_kernel<<<dim3(blocks, blocks), 64, 0, stream>>>(n, thr, d_boxes, mask, rotated);
  MODEST_LAUNCH_CHECK("nms_mask_kernel");
  nms_reduce_kernel<<<1, 32, 0, stream>>>(n, mask, keep, num_dev, removed);
  MODEST_LAUNCH_CHECK("nms_reduce_kernel");
  note_launch(2);
  // legacy contract (iou3d_nms.cpp:90-136): the count is returned to the host, so this call
  // synchronises the stream
  MODEST_CUDA(cudaMemcpyAsync(h_num, num_dev, sizeof(int), cudaMemcpyDeviceToHost, stream));
  MODEST_CUDA(cudaStreamSynchronize(stream));
  if (h_keep && *h_num > 0) {
    MODEST_CUDA(cudaMemcpyAsync(h_keep, keep, sizeof(long long) * (size_t)*h_num, cudaMemcpyDeviceToHost, stream));
    MODEST_CUDA(cudaStreamSynchronize(stream));
  }
  return MODEST_OK;
}

extern "C" int modest_nms_bev(const float* d_boxes, int n, float thresh, int64_t* d_keep, int64_t* h_keep,
                              int* h_num_out, void* d_ws, size_t ws_bytes, void* stream) {
  return nms_impl(d_boxes, n, thresh, reinterpret_cast<long long*>(d_keep), reinterpret_cast<long long*>(h_keep),
                  h_num_out, d_ws, ws_bytes, static_cast<cudaStream_t>(stream), 1);
}
extern "C" int modest_nms_normal(const float* d_boxes, int n, float thresh, int64_t* d_keep, int64_t* h_keep,
                                 int* h_num_out, void* d_ws, size_t ws_bytes, void* stream) {
  return nms_impl(d_boxes, n, thresh, reinterpret_cast<long long*>(d_keep), reinterpret_cast<long long*>(h_keep),
                  h_num_out, d_ws, ws_bytes, static_cast<cudaStream_t>(stream), 0);
}

extern "C" int modest_seed_nms_batch(const double* d_boxes, const int32_t* d_n_boxes, int n_scans, int max_boxes,
                                     float thresh, float* d_iou_or_null, float* d_iou_ws, uint8_t* d_keep, void* stream_) {
  modest::StageRange nvtx_("modest:N seed NMS");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_scans <= 0) return MODEST_OK;
  MODEST_REQUIRE(d_boxes && d_n_boxes && d_iou_ws && d_keep, "seed_nms: null pointer argument");
  MODEST_REQUIRE(max_boxes >= 1 && max_boxes <= 1024, "seed_nms: max_boxes %d out of range", max_boxes);
  const size_t smem = sizeof(float) * 7 * max_boxes + sizeof(int) * max_boxes;
  seed_nms_kernel<<<n_scans, 256, smem, stream>>>(d_boxes, d_n_boxes, max_boxes, thresh, d_iou_or_null, d_iou_ws, d_keep);
  MODEST_LAUNCH_CHECK("seed_nms_kernel");
  note_launch(1);
  return MODEST_OK;
}
