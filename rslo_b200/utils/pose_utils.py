"""Quaternion algebra used on the hot path (`rslo/utils/pose_utils.py:23-180`; quaternions are
(w, x, y, z) here) and the two quaternion<->matrix conversions the reference takes from kornia 0.4.0
(x, y, z, w order; call sites `voxel_odom_net.py:675-676,729-731`, `losses.py:359`)."""
import torch
import torch.nn.functional as F


def qinv(q):
    return torch.cat((q[:, :1], -q[:, 1:]), dim=1)


def qmult(q1, q2):
    q1s, q1v = q1[:, :1], q1[:, 1:]
    q2s, q2v = q2[:, :1], q2[:, 1:]
    qs = q1s * q2s - (q1v * q2v).sum(1, keepdim=True)
    qv = q1v * q2s + q2v * q1s + torch.cross(q1v, q2v, dim=1)
    q = torch.cat((qs, qv), dim=1)
    return q / q.norm(p=2, dim=1, keepdim=True)


def rotate_vec_by_q(t, q):
    """t' = t + 2 qs (qv x t) + 2 qv x (qv x t)."""
    qs, qv = q[:, :1], q[:, 1:]
    b = torch.cross(qv, t, dim=1)
    c = 2 * torch.cross(qv, b, dim=1)
    b = 2 * b * qs
    return t + b + c


def quaternion_to_rotation_matrix(quaternion):
    """(x,y,z,w) -> [N,3,3]; normalises first with eps 1e-12 (kornia 0.4.0 semantics)."""
    q = F.normalize(quaternion, p=2, dim=-1, eps=1e-12)
    x, y, z, w = q.unbind(-1)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    m = torch.stack([1.0 - (tyy + tzz), txy - twz, txz + twy,
                     txy + twz, 1.0 - (txx + tzz), tyz - twx,
                     txz - twy, tyz + twx, 1.0 - (txx + tyy)], dim=-1)
    return m.view(-1, 3, 3)


def rotation_matrix_to_quaternion(R, eps=1e-8):
    """[N,3,3] -> (x,y,z,w), four-branch trace method (kornia 0.4.0 semantics)."""
    tiny = torch.finfo(R.dtype).tiny

    def sdiv(a, b):
        return a / torch.clamp(b, min=tiny)

    v = R.reshape(*R.shape[:-2], 9)
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.chunk(v, 9, dim=-1)
    trace = m00 + m11 + m22
    sq0 = torch.sqrt(trace + 1.0) * 2.0
    b0 = torch.cat([sdiv(m21 - m12, sq0), sdiv(m02 - m20, sq0), sdiv(m10 - m01, sq0), 0.25 * sq0], -1)
    sq1 = torch.sqrt(1.0 + m00 - m11 - m22 + eps) * 2.0
    b1 = torch.cat([0.25 * sq1, sdiv(m01 + m10, sq1), sdiv(m02 + m20, sq1), sdiv(m21 - m12, sq1)], -1)
    sq2 = torch.sqrt(1.0 + m11 - m00 - m22 + eps) * 2.0
    b2 = torch.cat([sdiv(m01 + m10, sq2), 0.25 * sq2, sdiv(m12 + m21, sq2), sdiv(m02 - m20, sq2)], -1)
    sq3 = torch.sqrt(1.0 + m22 - m00 - m11 + eps) * 2.0
    b3 = torch.cat([sdiv(m02 + m20, sq3), sdiv(m12 + m21, sq3), 0.25 * sq3, sdiv(m10 - m01, sq3)], -1)
    w2 = torch.where(m11 > m22, b2, b3)
    w1 = torch.where((m00 > m11) & (m00 > m22), b1, w2)
    return torch.where(trace > 0.0, b0, w1)
