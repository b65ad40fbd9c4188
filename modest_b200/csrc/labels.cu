// Stage O (host side): field-of-view gate and KITTI label text.
//
// Reference behaviour: is_within_fov() and objs2label() (utils/pointcloud_utils.py:347-379)
// with compute_box_3d()/project_to_image() (utils/kitti_util.py:383-389,405-478).
// Pure host arithmetic on a handful of boxes per scan; lives in the library so that the
// batched pipeline can turn device results into label blobs without per-scan Python work.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace modest {

static void project(const double* P, double x, double y, double z, double* u, double* v) {
  // [x y z 1] @ P^T, then divide by the third component
  // numpy's (n,4) @ (4,3) goes through dgemm: k-ordered fused multiply-adds
  const double a = fma(1.0, P[3], fma(z, P[2], fma(y, P[1], x * P[0])));
  const double b = fma(1.0, P[7], fma(z, P[6], fma(y, P[5], x * P[4])));
  const double c = fma(1.0, P[11], fma(z, P[10], fma(y, P[9], x * P[8])));
  *u = a / c;
  *v = b / c;
}

}  // namespace modest

using namespace modest;

// h_boxes: (n,8) f64 rows [t.x, t.y, t.z, l, w, h, ry, volume]; h_P: 3x4 row-major P2.
// h_keep_in: optional (n) u8 pre-mask (e.g. the NMS result).  Writes the surviving boxes'
// label lines ('\n'-joined, no trailing newline, NUL-terminated) into h_text and returns
// the text length in *h_len; *h_n_out = number of lines.  h_kept_out (n) u8 optional.
extern "C" int modest_kitti_labels_host(const double* h_boxes, int n, const uint8_t* h_keep_in, const double* h_P,
                                        int fov_only, int image_h, int image_w, const char* obj_type,
                                        const double* h_scores, char* h_text, size_t text_cap, size_t* h_len,
                                        int* h_n_out, uint8_t* h_kept_out) {
  MODEST_REQUIRE(h_P && h_text && h_len && h_n_out && (n == 0 || h_boxes), "kitti_labels: null pointer argument");
  MODEST_REQUIRE(text_cap >= 1, "kitti_labels: empty text buffer");
  size_t pos = 0;
  int lines = 0;
  h_text[0] = 0;
  const char* type = obj_type ? obj_type : "Dynamic";
  for (int k = 0; k < n; ++k) {
    if (h_kept_out) h_kept_out[k] = 0;
    if (h_keep_in && !h_keep_in[k]) continue;
    const double* b = h_boxes + 8 * (size_t)k;
    const double tx = b[0], ty = b[1], tz = b[2], l = b[3], w = b[4], h = b[5], ry = b[6];
    if (fov_only) {
      // centre = t - [0, h/2, 0]; inside the image and in front of the camera
      double u, v;
      const double cy = ty - h / 2;
      project(h_P, tx, cy, tz, &u, &v);
      if (!(u < (double)image_w && u >= 0 && v < (double)image_h && v >= 0 && tz > 0)) continue;
    }
    const double alpha = -atan2(tx, tz) + ry;
    const double c = cos(ry), s = sin(ry);
    const double xc[8] = {l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2};
    const double yc[8] = {0, 0, 0, 0, -h, -h, -h, -h};
    const double zc[8] = {w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2};
    double umin = INFINITY, vmin = INFINITY, umax = -INFINITY, vmax = -INFINITY;
    for (int q = 0; q < 8; ++q) {
      // roty(ry) @ [x;y;z] + t
      const double X = fma(s, zc[q], fma(0.0, yc[q], c * xc[q])) + tx;
      const double Y = fma(0.0, zc[q], fma(1.0, yc[q], 0.0 * xc[q])) + ty;
      const double Z = fma(c, zc[q], fma(0.0, yc[q], -s * xc[q])) + tz;
      double u, v;
      project(h_P, X, Y, Z, &u, &v);
      umin = fmin(umin, u); umax = fmax(umax, u); vmin = fmin(vmin, v); vmax = fmax(vmax, v);
    }
    char line[512];
    int len;
    if (h_scores)
      len = snprintf(line, sizeof(line), "%s -1 -1 %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f",
                     type, alpha, umin, vmin, umax, vmax, h, w, l, tx, ty, tz, ry, h_scores[k]);
    else
      len = snprintf(line, sizeof(line), "%s -1 -1 %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f",
                     type, alpha, umin, vmin, umax, vmax, h, w, l, tx, ty, tz, ry);
    if (len < 0) len = 0;
    const size_t need = (size_t)len + (lines ? 1 : 0);
    if (pos + need + 1 > text_cap) {
      set_error("kitti_labels: text buffer too small (%zu bytes)", text_cap);
      return MODEST_ERR_CAPACITY;
    }
    if (lines) h_text[pos++] = '\n';
    memcpy(h_text + pos, line, (size_t)len);
    pos += (size_t)len;
    h_text[pos] = 0;
    ++lines;
    if (h_kept_out) h_kept_out[k] = 1;
  }
  *h_len = pos;
  *h_n_out = lines;
  return MODEST_OK;
}

// Batched form: scan s has h_n_boxes[s] rows at h_boxes + s*max_boxes*8, keep flags at
// h_keep + s*max_boxes (optional) and P2 at h_P + s*12.  Texts are written back to back into
// h_text (no separators); h_text_off (n_scans+1) receives their byte offsets.
extern "C" int modest_kitti_labels_batch_host(const double* h_boxes, const int32_t* h_n_boxes, const uint8_t* h_keep,
                                              int n_scans, int max_boxes, const double* h_P, int fov_only, int image_h,
                                              int image_w, const char* obj_type, char* h_text, size_t text_cap,
                                              int64_t* h_text_off) {
  MODEST_REQUIRE(h_n_boxes && h_P && h_text && h_text_off && (n_scans == 0 || h_boxes), "kitti_labels_batch: null pointer argument");
  size_t pos = 0;
  h_text_off[0] = 0;
  for (int s = 0; s < n_scans; ++s) {
    size_t len = 0;
    int lines = 0;
    MODEST_REQUIRE(pos < text_cap, "kitti_labels_batch: text buffer too small");
    const int rc = modest_kitti_labels_host(h_boxes + (size_t)s * max_boxes * 8, h_n_boxes[s],
                                            h_keep ? h_keep + (size_t)s * max_boxes : nullptr, h_P + (size_t)s * 12, fov_only,
                                            image_h, image_w, obj_type, nullptr, h_text + pos, text_cap - pos, &len, &lines,
                                            nullptr);
    if (rc != MODEST_OK) return rc;
    pos += len;
    h_text_off[s + 1] = (int64_t)pos;
  }
  return MODEST_OK;
}
