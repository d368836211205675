"""Second, independent restatement (numpy, matrix form) of the reference's forward kinematics
and energy, used only to cross-check oracle/hpe_oracle.c.  Written from the DH matrices in
src/fingermodel.cpp:106-180, src/thumbmodel.cpp:112-193 and the chain orders at
src/fingermodel.cpp:285-311, src/thumbmodel.cpp:287-313 (paths relative to /root/reference).
"""
import numpy as np

PI = np.arccos(-1.0)


def _dh(a, L):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0, L * c], [s, c, 0, L * s], [0, 0, 1, 0], [0, 0, 0, 1.0]])


def _tw(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, -s, 0], [s, 0, c, 0], [0, -1, 0, 0], [0, 0, 0, 1.0]])


def _rad(d):
    return d / 180.0 * PI


def _global(theta):
    z, y, x = _rad(theta[0] + 180), _rad(theta[1]), _rad(theta[2])
    Rz = np.array([[np.cos(z), -np.sin(z), 0, 0], [np.sin(z), np.cos(z), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    Ry = np.array([[np.cos(y), 0, np.sin(y), 0], [0, 1, 0, 0], [-np.sin(y), 0, np.cos(y), 0], [0, 0, 0, 1.0]])
    Rx = np.array([[1, 0, 0, 0], [0, np.cos(x), -np.sin(x), 0], [0, np.sin(x), np.cos(x), 0], [0, 0, 0, 1.0]])
    T = np.eye(4)
    T[:3, 3] = theta[3:6]
    return T @ (Rz @ Ry @ Rx)


def digit_joints(d, theta, geo, spacing, cmc):
    L = geo[4 * d:4 * d + 4]
    sp = float(np.float32(spacing[d]))
    C = _rad(cmc[d])
    a = np.sqrt(L[0] ** 2 + float(np.float32(sp) * np.float32(sp)) - 2 * L[0] * sp * np.cos(C))
    beta = np.arcsin(np.sin(C) * sp / a)
    base = _dh(C, L[0])
    back = np.eye(4)
    back[:2, :2] = [[np.cos(beta), -np.sin(beta)], [np.sin(beta), np.cos(beta)]]
    k = a if d == 0 else L[0] * np.sin(C)
    back[0, 3], back[1, 3] = -k * np.cos(beta), -k * np.sin(beta)
    th = [_rad(t) for t in theta[6 + 4 * d:10 + 4 * d]]
    if d == 0:
        p = C + PI
        c2, s2 = np.cos(th[1]), np.sin(th[1])
        T12 = np.array([[c2, -s2 * np.cos(p), s2 * np.sin(p), L[1] * c2],
                        [s2, c2 * np.cos(p), -c2 * np.sin(p), L[1] * s2],
                        [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1.0]])
        second = _tw(th[0]) @ T12
    else:
        second = _tw(th[0]) @ _dh(th[1], L[1])
    cur = _global(theta) @ base
    J = np.zeros((5, 3))
    J[1] = cur[:3, 3]
    J[0] = (cur @ back)[:3, 3]
    cur = cur @ second
    J[2] = cur[:3, 3]
    cur = cur @ _dh(th[2], L[2])
    J[3] = cur[:3, 3]
    cur = cur @ _dh(th[3], L[3])
    J[4] = cur[:3, 3]
    return J


def hand(theta, geo, spacing, cmc, tb_n=(2, 2, 2, 2), fg_n=(4, 2, 2, 2)):
    S = []
    joints = np.zeros((21, 3))
    joints[0] = theta[3:6]
    for d in range(5):
        J = digit_joints(d, theta, geo, spacing, cmc)
        n = tb_n if d == 0 else fg_n
        for i in range(4):
            if d > 0 and i == 0:
                ts = [j / (n[i] - 1) for j in range(n[i])]
            else:
                ts = [j / n[i] for j in range(1, n[i] + 1)]
            for t in ts:
                S.append((1 - t) * J[i] + t * J[i + 1])
        if d == 0:
            joints[17:21] = J[1:]
        else:
            joints[1 + 4 * (d - 1):5 + 4 * (d - 1)] = J[1:]
    S = np.array(S)
    S[:, 1:] *= -1
    return S, joints


def cost_terms(theta, geo, spacing, cmc, rad, cloud, depth_cm, dt_px, scale, focal=241.42):
    """align, depth, collision following src/costfunc.cpp:130-197, 227-304, 306-377."""
    S, _ = hand(theta, geo, spacing, cmc)
    P32, S32 = cloud.astype(np.float32), S.astype(np.float32)
    t = P32[:, None, :] - S32[None, :, :]
    sq = t * t
    d = np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])
    m = d.argmin(1)
    nd = np.sqrt(((cloud - S[m]) ** 2).sum(1))
    align = ((nd - rad[m]) ** 2).sum() * (48.0 / cloud.shape[0])
    U = S.copy()
    U[:, 1:] *= -1
    K = np.array([[focal, 0, 160.0], [0, focal, 120.0], [0, 0, 1.0]])
    pr = K @ U.T
    pr = np.floor(pr[:2] / pr[2])
    depth = 0.0
    for i in range(48):
        dx, dy = pr[0, i], pr[1, i]
        if 0 <= dx < 320 and 0 <= dy < 240:
            dj = depth_cm[int(dy), int(dx)]
            if dj != 0:
                depth += max(0.0, dj - U[i, 2]) ** 2
            else:
                depth += (dt_px[int(dy), int(dx)] * scale + rad[i]) ** 2
        else:
            depth += (dt_px.max() * scale + rad[i]) ** 2
    first = [2, 12, 22, 32, 42]
    col = 0.0
    for i in range(4):
        A, B = U[first[i]:first[i] + 6], U[first[i + 1]:first[i + 1] + 6]
        rA, rB = rad[first[i]:first[i] + 6], rad[first[i + 1]:first[i + 1] + 6]
        dist = np.sqrt(((A[:, None, :] - B[None, :, :]) ** 2).sum(-1))
        pen = rA[:, None] + rB[None, :] - dist
        col += (pen[pen > 0] ** 2).sum()
    return align, depth, col
