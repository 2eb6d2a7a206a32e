"""Host-side geometry helpers with the reference's numerics (reference math.py:6-62).

`intersect_circle_segs` is the numpy form the device kernel (csrc/physics.cuh) is
checked against: per segment fl(fl(ap.x*ab.x)+fl(ap.z*ab.z)) etc. -- ufunc sums, no FMA.
"""
import math

import numpy as np

X_VEC = np.array([1, 0, 0])
Y_VEC = np.array([0, 1, 0])
Z_VEC = np.array([0, 0, 1])


def gen_rot_matrix(axis, angle):
    """Counter-clockwise rotation about `axis` by `angle` radians (quaternion form)."""
    axis = axis / math.sqrt(np.dot(axis, axis))
    a = math.cos(angle / 2.0)
    b, c, d = -axis * math.sin(angle / 2.0)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    return np.array([
        [aa + bb - cc - dd, 2 * (b * c - a * d), 2 * (b * d + a * c)],
        [2 * (b * c + a * d), aa + cc - bb - dd, 2 * (c * d - a * b)],
        [2 * (b * d - a * c), 2 * (c * d + a * b), aa + dd - bb - cc],
    ])


def intersect_circle_segs(point, radius, segs):
    """True if the xz-circle (point, radius) touches any segment of segs[S,2,3]; else None."""
    p = np.array([point[0], 0, point[2]])
    a, b = segs[:, 0, :], segs[:, 1, :]
    ab, ap = b - a, p - a
    t = np.sum(ap * ab, axis=1) / np.sum(ab * ab, axis=1)
    t = np.expand_dims(np.clip(t, 0, 1), axis=1)
    closest = a + t * ab
    dist = np.linalg.norm(closest - p, axis=1)
    return True if np.any(np.less(dist, radius)) else None
