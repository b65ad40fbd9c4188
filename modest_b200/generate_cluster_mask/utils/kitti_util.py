"""Drop-in for the parts of the reference's utils/kitti_util.py the hot path reaches
(SURVEY.md 8(a)-K/O): Calibration (calib txt parsing, velo->rect, rect->image) and
compute_box_3d / project_to_image / roty.  Host-side bookkeeping on a few 3x4 matrices; the
per-point projections of the pipeline itself run inside libmodest_b200 (csrc/boxes.cu)."""
import numpy as np


def inverse_rigid_trans(Tr):
    """utils/kitti_util.py (inverse of a 3x4 [R|t])"""
    inv = np.zeros_like(Tr)
    inv[0:3, 0:3] = np.transpose(Tr[0:3, 0:3])
    inv[0:3, 3] = np.dot(-np.transpose(Tr[0:3, 0:3]), Tr[0:3, 3])
    return inv


class Calibration(object):
    """utils/kitti_util.py:200-342 -- P2, Tr_velo_to_cam, R0_rect (+P3 for the baseline) parsed
    from a KITTI calibration file; `from_video` layout is not used on this path."""

    def __init__(self, calib_filepath, from_video=False):
        if from_video:
            raise NotImplementedError("from_video calibration is not on the seed-label path")
        calibs = calib_filepath if isinstance(calib_filepath, dict) else self.read_calib_file(calib_filepath)
        self.P = np.reshape(np.asarray(calibs['P2'], dtype=np.float64), [3, 4])
        self.V2C = np.reshape(np.asarray(calibs['Tr_velo_to_cam'], dtype=np.float64), [3, 4])
        self.C2V = inverse_rigid_trans(self.V2C)
        self.R0 = np.reshape(np.asarray(calibs['R0_rect'], dtype=np.float64), [3, 3])
        self.P3 = np.reshape(np.asarray(calibs.get('P3', np.zeros(12)), dtype=np.float64), [3, 4])
        self.c_u, self.c_v = self.P[0, 2], self.P[1, 2]
        self.f_u, self.f_v = self.P[0, 0], self.P[1, 1]
        self.b_x = self.P[0, 3] / (-self.f_u)
        self.b_y = self.P[1, 3] / (-self.f_v)
        self.baseline = self.P3[0, 3] / (-self.f_u) - self.P[0, 3] / (-self.f_u)

    @staticmethod
    def read_calib_file(filepath):
        data = {}
        with open(filepath, 'r') as f:
            for line in f.readlines():
                line = line.rstrip()
                if len(line) == 0:
                    continue
                key, value = line.split(':', 1)
                try:
                    data[key] = np.array([float(x) for x in value.split()])
                except ValueError:
                    pass
        return data

    def cart2hom(self, pts_3d):
        return np.hstack((pts_3d, np.ones((pts_3d.shape[0], 1))))

    def project_velo_to_ref(self, pts_3d_velo):
        return np.dot(self.cart2hom(pts_3d_velo), np.transpose(self.V2C))

    def project_ref_to_rect(self, pts_3d_ref):
        return np.transpose(np.dot(self.R0, np.transpose(pts_3d_ref)))

    def project_velo_to_rect(self, pts_3d_velo):
        return self.project_ref_to_rect(self.project_velo_to_ref(pts_3d_velo))

    def project_rect_to_image(self, pts_3d_rect):
        pts_2d = np.dot(self.cart2hom(pts_3d_rect), np.transpose(self.P))
        pts_2d[:, 0] /= pts_2d[:, 2]
        pts_2d[:, 1] /= pts_2d[:, 2]
        return pts_2d[:, 0:2]

    def project_velo_to_image(self, pts_3d_velo):
        return self.project_rect_to_image(self.project_velo_to_rect(pts_3d_velo))


def roty(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def project_to_image(pts_3d, P):
    pts_2d = np.dot(np.hstack((pts_3d, np.ones((pts_3d.shape[0], 1)))), np.transpose(P))
    pts_2d[:, 0] /= pts_2d[:, 2]
    pts_2d[:, 1] /= pts_2d[:, 2]
    return pts_2d[:, 0:2]


def compute_box_3d(obj, P):
    """utils/kitti_util.py:430-478 -- (8,2) image corners and (8,3) rect corners of a box."""
    hl, hw, h = obj.l / 2, obj.w / 2, obj.h
    local = np.vstack([[hl, hl, -hl, -hl, hl, hl, -hl, -hl], [0, 0, 0, 0, -h, -h, -h, -h],
                       [hw, -hw, -hw, hw, hw, -hw, -hw, hw]])
    corners_3d = np.dot(roty(obj.ry), local)
    corners_3d[0, :] += obj.t[0]
    corners_3d[1, :] += obj.t[1]
    corners_3d[2, :] += obj.t[2]
    return project_to_image(np.transpose(corners_3d), P), np.transpose(corners_3d)
