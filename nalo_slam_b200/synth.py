"""Seeded synthetic inputs for the photometric-alignment path (SURVEY.md §8(d) "Synthetic inputs").

The reference ships no data and there is no network, so every test and benchmark runs on analytic scenes:

* texture  I(x,y) = clip(127.5 + sum_k A_k sin(2*pi*(fx_k x + fy_k y) + phi_k), 0, 255), 24 sinusoids,
  spatial periods 6..200 px (smooth => wide Gauss-Newton basin, gradients everywhere);
* inverse depth in the reference frame: tilted plane + smooth bumps, clipped to [0.02, 0.5];
* KITTI-like intrinsics fx=fy=718.856, cx=607.19, cy=185.22 at 1241x376 (scaled for smaller test images);
* the new frame is rendered by *inverse* warping the analytic texture (fixed-point inversion of the forward
  warp, no resampling blur => the residual floor is ~0) and applying the affine brightness change.

Nothing here depends on the oracle or on the CUDA extension; it is plain numpy.
Conventions follow the reference: pose = refToNew, p_new = R p_ref + t; SE3 stored as
[qx,qy,qz,qw,tx,ty,tz] (Sophus/Eigen order); tangent = [v(3), omega(3)] (thirdparty/Sophus/sophus/se3.hpp:407-428).
"""
from __future__ import annotations

import dataclasses

import numpy as np

KITTI_W, KITTI_H = 1241, 376
KITTI_K = (718.856, 718.856, 607.19, 185.22)
DEFAULT_SEED = 20261018


def pyr_sizes(w0: int, h0: int, levels: int):
    return [(w0 >> l, h0 >> l) for l in range(levels)]


def scaled_K(w: int, h: int):
    """KITTI intrinsics scaled to a w x h image (for the small CPU-test sizes)."""
    s = w / KITTI_W
    fx, fy, cx, cy = KITTI_K
    return (fx * s, fy * s, (cx + 0.5) * s - 0.5, (cy + 0.5) * (h / KITTI_H) - 0.5)


# ---------------------------------------------------------------- SE3 (host-side convenience, numpy/double)
def so3_exp_quat(omega):
    omega = np.asarray(omega, dtype=np.float64)
    th = float(np.linalg.norm(omega))
    if th < 1e-10:
        imag = 0.5 - th * th / 48.0
        real = 1.0 - th * th / 8.0
    else:
        imag = np.sin(0.5 * th) / th
        real = np.cos(0.5 * th)
    q = np.array([imag * omega[0], imag * omega[1], imag * omega[2], real])
    return q / np.linalg.norm(q)


def quat_to_R(q):
    x, y, z, w = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def se3_exp(xi):
    """tangent [v, omega] -> pose7 [qx,qy,qz,qw,tx,ty,tz]."""
    xi = np.asarray(xi, dtype=np.float64)
    v, om = xi[:3], xi[3:]
    th = float(np.linalg.norm(om))
    q = so3_exp_quat(om)
    O = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    if th < 1e-10:
        V = quat_to_R(q)
    else:
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * O + (th - np.sin(th)) / th**3 * (O @ O)
    return np.concatenate([q, V @ v])


def pose_identity():
    return np.array([0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


def pose_distance(p, q):
    """(translation error [m], rotation error [rad]) between two pose7."""
    dt = float(np.linalg.norm(p[4:] - q[4:]))
    Ra, Rb = quat_to_R(p[:4]), quat_to_R(q[:4])
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return dt, float(np.arccos(np.clip(c, -1, 1)))


# ---------------------------------------------------------------- scene
@dataclasses.dataclass
class Scene:
    w: int
    h: int
    K: tuple
    amp: np.ndarray
    fxk: np.ndarray
    fyk: np.ndarray
    phi: np.ndarray
    plane: np.ndarray  # idepth = p0 + p1*xn + p2*yn + bumps
    bumps: np.ndarray  # rows of (amp, cx, cy, sigma) in normalised coords

    def texture(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        acc = np.full(x.shape, 127.5)
        for a, fx, fy, ph in zip(self.amp, self.fxk, self.fyk, self.phi):
            acc += a * np.sin(2 * np.pi * (fx * x + fy * y) + ph)
        return np.clip(acc, 0.0, 255.0)

    def idepth(self, x, y):
        xn = np.asarray(x, dtype=np.float64) / self.w - 0.5
        yn = np.asarray(y, dtype=np.float64) / self.h - 0.5
        d = self.plane[0] + self.plane[1] * xn + self.plane[2] * yn
        for a, cx, cy, s in self.bumps:
            d = d + a * np.exp(-((xn - cx) ** 2 + (yn - cy) ** 2) / (2 * s * s))
        return np.clip(d, 0.02, 0.5)


def make_scene(w=KITTI_W, h=KITTI_H, K=None, seed=DEFAULT_SEED, n_sin=24) -> Scene:
    rng = np.random.default_rng(seed)
    if K is None:
        K = KITTI_K if (w, h) == (KITTI_W, KITTI_H) else scaled_K(w, h)
    amp = rng.uniform(2.0, 12.0, n_sin)
    period = np.exp(rng.uniform(np.log(6.0), np.log(200.0), n_sin))
    ang = rng.uniform(0, 2 * np.pi, n_sin)
    fxk = np.cos(ang) / period
    fyk = np.sin(ang) / period
    phi = rng.uniform(0, 2 * np.pi, n_sin)
    plane = np.array([rng.uniform(0.12, 0.2), rng.uniform(-0.08, 0.08), rng.uniform(0.05, 0.15)])
    bumps = np.stack(
        [rng.uniform(-0.05, 0.05, 4), rng.uniform(-0.4, 0.4, 4), rng.uniform(-0.4, 0.4, 4), rng.uniform(0.08, 0.25, 4)], axis=1
    )
    return Scene(w, h, tuple(K), amp, fxk, fyk, phi, plane, bumps)


def random_motion(rng, scale=1.0):
    """xi ~ N(0, diag(sigma_t=(0.02,0.02,0.08) m, sigma_w=0.003 rad)), affine a in +-0.05, b in +-3."""
    xi = rng.normal(0, 1, 6) * np.array([0.02, 0.02, 0.08, 0.003, 0.003, 0.003]) * scale
    aff = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-3, 3)]) * scale
    return xi, aff


def forward_warp(scene: Scene, pose7, x, y):
    """ref pixel (x,y) -> new-frame pixel, using the scene's analytic inverse depth."""
    fx, fy, cx, cy = scene.K
    R = quat_to_R(pose7[:4])
    t = pose7[4:]
    idp = scene.idepth(x, y)
    X = (x - cx) / fx
    Y = (y - cy) / fy
    px = R[0, 0] * X + R[0, 1] * Y + R[0, 2] + t[0] * idp
    py = R[1, 0] * X + R[1, 1] * Y + R[1, 2] + t[1] * idp
    pz = R[2, 0] * X + R[2, 1] * Y + R[2, 2] + t[2] * idp
    return fx * px / pz + cx, fy * py / pz + cy


def render_ref(scene: Scene) -> np.ndarray:
    yy, xx = np.mgrid[0 : scene.h, 0 : scene.w]
    return scene.texture(xx, yy).astype(np.float32)


def render_new(scene: Scene, pose7, aff=(0.0, 0.0), iters=12) -> np.ndarray:
    """I_new(p') = exp(a) * I_ref(W^-1(p')) + b, W^-1 by fixed-point iteration (flow is small and smooth)."""
    yy, xx = np.mgrid[0 : scene.h, 0 : scene.w]
    tx = xx.astype(np.float64)
    ty = yy.astype(np.float64)
    sx, sy = tx.copy(), ty.copy()
    for _ in range(iters):
        wx, wy = forward_warp(scene, pose7, sx, sy)
        sx -= wx - tx
        sy -= wy - ty
    img = np.exp(aff[0]) * scene.texture(sx, sy) + aff[1]
    return img.astype(np.float32)


def dense_reference_maps(scene: Scene, absgrad0: np.ndarray, keep_fraction=0.43):
    """North-star dense mode: every L0 pixel whose squared gradient exceeds the
    (1-keep_fraction) quantile gets its ground-truth inverse depth with weight 1.
    Returns (idw0, wsum0) = level-0 inputs of makeCoarseDepthL0 step 2."""
    ag = absgrad0.reshape(scene.h, scene.w)
    tau = np.quantile(ag, 1.0 - keep_fraction)
    sel = ag > tau
    yy, xx = np.mgrid[0 : scene.h, 0 : scene.w]
    idp = scene.idepth(xx, yy).astype(np.float32)
    wsum = sel.astype(np.float32)
    idw = np.where(sel, idp, 0).astype(np.float32)
    return idw, wsum


def sparse_reference_points(scene: Scene, sel_map: np.ndarray, hdi=1e-3):
    """Sparse list for makeCoarseDepthL0 step 1 from a selection map (non-zero = selected):
    (u, v, idepth, HdiF) with ground-truth inverse depth."""
    ys, xs = np.nonzero(sel_map.reshape(scene.h, scene.w))
    u = xs.astype(np.float32)
    v = ys.astype(np.float32)
    idp = scene.idepth(xs, ys).astype(np.float32)
    return u, v, idp, np.full(u.shape, hdi, dtype=np.float32)


# ---------------------------------------------------------------- windowed-BA synthetic residual table
BA_REC_WORDS = 76
BA_O = dict(res=0, jpdxi=8, jpdc=20, jpdd=28, jidx=30, jab=46, jidx2=62, jabjidx=65, jab2=69, pt=72, pack=73)


def make_ba_problem(nf=7, pts_per_frame=100, seed=DEFAULT_SEED, lin_fraction=0.0, drop_fraction=0.05):
    """Flattened windowed-BA input (include/nalo_gpu.h NALO_BA_*): every point hosted in frame h has one
    residual to each other frame (config 4 of BASELINE.json). Jacobian magnitudes mimic
    PointFrameResidual::linearize (src/FullSystem/Residuals.cpp:78-274): image gradients O(10),
    d(x,y)/d(xi) O(fx*idepth), shorthand products computed from JIdx/JabF exactly as linearize does.
    Records are ordered by (host,target) bucket; pt_res keeps residualsAll order per point."""
    rng = np.random.default_rng(seed)
    n_pts = nf * pts_per_frame
    recs = []
    pt_lists = [[] for _ in range(n_pts)]
    # build in bucket order
    for t in range(nf):
        for h in range(nf):
            if h == t:
                continue
            for k in range(pts_per_frame):
                p = h * pts_per_frame + k
                if rng.random() < drop_fraction:
                    continue
                recs.append((h, t, p))
    n_res = len(recs)
    rec = np.zeros((n_res, BA_REC_WORDS), dtype=np.float32)
    rec_i = rec.view(np.int32)
    res = rng.normal(0, 4.0, (n_res, 8)).astype(np.float32)
    jidx = rng.normal(0, 8.0, (n_res, 2, 8)).astype(np.float32)
    hw = rng.uniform(0.3, 1.0, (n_res, 8)).astype(np.float32)
    drdA = rng.normal(0, 30.0, (n_res, 8)).astype(np.float32)
    jidx = (jidx * hw[:, None, :]).astype(np.float32)
    jab = np.stack([drdA * hw, hw], axis=1).astype(np.float32)
    rec[:, 0:8] = res * hw
    rec[:, 8:20] = rng.normal(0, 60.0, (n_res, 12)).astype(np.float32)
    rec[:, 20:28] = rng.normal(0, 0.3, (n_res, 8)).astype(np.float32)
    rec[:, 28:30] = rng.normal(0, 40.0, (n_res, 2)).astype(np.float32)
    rec[:, 30:46] = jidx.reshape(n_res, 16)
    rec[:, 46:62] = jab.reshape(n_res, 16)
    f32 = np.float32
    rec[:, 62] = np.sum(jidx[:, 0] * jidx[:, 0], axis=1, dtype=f32)
    rec[:, 63] = np.sum(jidx[:, 0] * jidx[:, 1], axis=1, dtype=f32)
    rec[:, 64] = np.sum(jidx[:, 1] * jidx[:, 1], axis=1, dtype=f32)
    rec[:, 65] = np.sum(jab[:, 0] * jidx[:, 0], axis=1, dtype=f32)
    rec[:, 66] = np.sum(jab[:, 0] * jidx[:, 1], axis=1, dtype=f32)
    rec[:, 67] = np.sum(jab[:, 1] * jidx[:, 0], axis=1, dtype=f32)
    rec[:, 68] = np.sum(jab[:, 1] * jidx[:, 1], axis=1, dtype=f32)
    rec[:, 69] = np.sum(jab[:, 0] * jab[:, 0], axis=1, dtype=f32)
    rec[:, 70] = np.sum(jab[:, 0] * jab[:, 1], axis=1, dtype=f32)
    rec[:, 71] = np.sum(jab[:, 1] * jab[:, 1], axis=1, dtype=f32)
    lin = rng.random(n_res) < lin_fraction
    inactive = rng.random(n_res) < 0.03
    for i, (h, t, p) in enumerate(recs):
        flags = (0 if inactive[i] else 1) | (2 if lin[i] else 0)
        rec_i[i, 72] = p
        rec_i[i, 73] = h | (t << 8) | (flags << 16)
        pt_lists[p].append(i)
    pt_begin = np.zeros(n_pts + 1, dtype=np.int32)
    for p in range(n_pts):
        pt_begin[p + 1] = pt_begin[p] + len(pt_lists[p])
    pt_res = np.array([i for lst in pt_lists for i in lst], dtype=np.int32)
    bucket_begin = np.zeros(nf * nf + 1, dtype=np.int32)
    for h, t, _ in recs:
        bucket_begin[h + t * nf + 1] += 1
    bucket_begin = np.cumsum(bucket_begin).astype(np.int32)
    # records were generated t-major/h-minor == ascending htIDX = h + t*nf, so they are bucket-sorted
    return dict(
        nf=nf,
        n_pts=n_pts,
        n_res=n_res,
        rec=rec,
        res_toZero=rng.normal(0, 4.0, (n_res, 8)).astype(np.float32),
        pt_begin=pt_begin,
        pt_res=pt_res,
        bucket_begin=bucket_begin,
        deltaF=rng.normal(0, 0.01, n_pts).astype(np.float32),
        priorF=np.zeros(n_pts, dtype=np.float32),
        adHTdeltaF=rng.normal(0, 1e-3, (nf * nf, 8)).astype(np.float32),
        cDeltaF=rng.normal(0, 1e-2, 4).astype(np.float32),
    )


# ---------------------------------------------------------------- linearize (f1) synthetic problem
LIN_PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], dtype=np.int32)  # staticPattern[8]
LIN_PAIR_WORDS = 32


def _pose_Rt(pose7):
    return quat_to_R(pose7[:4]), np.asarray(pose7[4:], dtype=np.float64)


def _bilinear(img, x, y):
    ix, iy = np.floor(x).astype(int), np.floor(y).astype(int)
    dx, dy = x - ix, y - iy
    return (img[iy, ix] * (1 - dx) * (1 - dy) + img[iy, ix + 1] * dx * (1 - dy) + img[iy + 1, ix] * (1 - dx) * dy + img[iy + 1, ix + 1] * dx * dy)


def make_lin_problem(scene: Scene, nf=4, pts_per_frame=500, seed=DEFAULT_SEED, bad_depth_fraction=0.05, fej_noise=1e-3):
    """Inputs of PointFrameResidual::linearize (src/FullSystem/Residuals.cpp:78-274) for a window of nf keyframes that all
    see the analytic scene: frame k is the reference view warped by a small random motion T_k (frame 0 = identity).
    Points are sampled in frame 0, moved into their host frame h (pixel and inverse depth), and get one residual to every
    other frame, flattened and bucket-sorted like the BA records (include/nalo_gpu.h). A few points get a wrong depth
    (-> OUTLIER) and some lie near the border (-> OOB), so all three result states occur.
    Returns a dict with the flat arrays, the nf images, and the per-(host,target) precalc table."""
    rng = np.random.default_rng(seed)
    w, h = scene.w, scene.h
    fx, fy, cx, cy = scene.K
    Kmat = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    Ki = np.linalg.inv(Kmat)
    poses, affs, imgs = [], [], []
    for k in range(nf):
        if k == 0:
            pose, aff = pose_identity(), np.zeros(2)
        else:
            xi, aff = random_motion(rng)
            pose = se3_exp(xi)
        poses.append(pose)
        affs.append(aff)
        imgs.append(render_new(scene, pose, aff) if k else render_ref(scene))
    # frame-frame precalc (HessianBlocks.cpp:192-222); the evaluation point differs slightly from the current state (FEJ)
    pairs = np.zeros((nf * nf, LIN_PAIR_WORDS), dtype=np.float32)
    rel = {}
    for hst in range(nf):
        Rh, th = _pose_Rt(poses[hst])
        for tgt in range(nf):
            Rt_, tt = _pose_Rt(poses[tgt])
            R = Rt_ @ Rh.T
            t = tt - R @ th
            rel[(hst, tgt)] = (R, t)
            om = rng.normal(0, fej_noise, 3)
            R0 = quat_to_R(so3_exp_quat(om)) @ R
            t0 = t + rng.normal(0, fej_noise * 0.1, 3)
            P = pairs[hst + tgt * nf]
            P[0:9] = R0.astype(np.float32).ravel()
            P[9:12] = t0
            P[12:21] = (Kmat @ R @ Ki).astype(np.float32).ravel()
            P[21:24] = Kmat @ t
            a = np.exp(affs[tgt][0] - affs[hst][0])
            P[24] = a
            P[25] = affs[tgt][1] - a * affs[hst][1]
            P[26] = affs[hst][1]
            P[27] = 8 * 12 * 12.0  # frameEnergyTH ~ patternNum * setting_outlierTH of a fresh frame (FullSystem.cpp setNewFrameEnergyTH)
            P.view(np.int32)[28] = tgt
    pt_id = 0
    per_bucket = {}
    pat = LIN_PATTERN.astype(np.float64)
    for hst in range(nf):
        x0 = rng.uniform(8, w - 9, pts_per_frame)
        y0 = rng.uniform(8, h - 9, pts_per_frame)
        border = rng.random(pts_per_frame) < 0.04
        x0[border] = rng.choice([2.5, w - 3.5], border.sum())
        id0 = scene.idepth(x0, y0)
        # into the host frame
        Rh, th = _pose_Rt(poses[hst])
        X = np.stack([(x0 - cx) / fx, (y0 - cy) / fy, np.ones_like(x0)])
        p = Rh @ X + th[:, None] * id0
        uh = fx * p[0] / p[2] + cx
        vh = fy * p[1] / p[2] + cy
        idh = id0 / p[2]
        bad = rng.random(pts_per_frame) < bad_depth_fraction
        idh_used = np.where(bad, idh * rng.uniform(1.5, 3.0, pts_per_frame), idh)
        inside = (uh > 3) & (uh < w - 4) & (vh > 3) & (vh < h - 4)
        uh, vh, idh_used = uh[inside], vh[inside], idh_used[inside]
        m = len(uh)
        img64 = imgs[hst].astype(np.float64)
        gyh, gxh = np.gradient(img64)
        px, py = uh[:, None] + pat[None, :, 0], vh[:, None] + pat[None, :, 1]
        cols = _bilinear(img64, px, py)
        g2 = _bilinear(gxh, px, py) ** 2 + _bilinear(gyh, px, py) ** 2
        wts = np.sqrt(2500.0 / (2500.0 + g2))
        ids = np.arange(pt_id, pt_id + m)
        pt_id += m
        for tgt in range(nf):
            if tgt == hst:
                continue
            idz = idh_used * (1 + rng.normal(0, fej_noise, m)) if fej_noise > 0 else idh_used.copy()
            per_bucket[hst + tgt * nf] = (np.stack([uh, vh, idz, idh_used], 1), cols, wts,
                                          np.full(m, hst | (tgt << 8) | (1 << 16), dtype=np.uint32), ids)
    keys = sorted(per_bucket)
    pt4 = np.concatenate([per_bucket[k][0] for k in keys]) if keys else np.zeros((0, 4))
    color = np.concatenate([per_bucket[k][1] for k in keys]) if keys else np.zeros((0, 8))
    weights = np.concatenate([per_bucket[k][2] for k in keys]) if keys else np.zeros((0, 8))
    pack = np.concatenate([per_bucket[k][3] for k in keys]) if keys else np.zeros(0, dtype=np.uint32)
    point = np.concatenate([per_bucket[k][4] for k in keys]) if keys else np.zeros(0, dtype=np.int64)
    n = len(pack)
    return dict(nf=nf, n_res=n, n_pts=pt_id, w=w, h=h, K=(fx, fy, cx, cy), images=imgs, pairs=pairs,
                pt4=np.array(pt4, dtype=np.float32).reshape(n, 4), color=np.array(color, dtype=np.float32).reshape(n, 8),
                weights=np.array(weights, dtype=np.float32).reshape(n, 8), pack=np.array(pack, dtype=np.uint32),
                point=np.array(point, dtype=np.int32), state_in=np.zeros(n, dtype=np.uint8), energy_in=np.zeros(n, dtype=np.float32))


# ---------------------------------------------------------------- CoarseInitializer point sets
def make_init_points(scene: Scene, lvl: int, step: int = 3, seed=DEFAULT_SEED, idepth_noise=0.05, bad_fraction=0.03, border=None):
    """Synthetic `Pnt` set of one level as CoarseInitializer::setFirst lays it out (:824-858): integer pixels + 0.1 inside
    the patternPadding border (here a regular grid every `step` px instead of the selector's picks), idepth_new = scaled
    ground truth with noise, a few points already marked bad, outlierTH = patternNum * setting_outlierTH."""
    rng = np.random.default_rng(seed + 17 * lvl)
    wl, hl = scene.w >> lvl, scene.h >> lvl
    pad = 2
    lo = pad + 1 if border is None else border
    xs = np.arange(lo, wl - pad - 2 if border is None else wl - border, step)
    ys = np.arange(lo, hl - pad - 2 if border is None else hl - border, step)
    xx, yy = np.meshgrid(xs, ys)
    xx, yy = xx.ravel(), yy.ravel()
    n = xx.size
    s = float(1 << lvl)
    idp = scene.idepth((xx + 0.5) * s - 0.5, (yy + 0.5) * s - 0.5)
    idp = idp / np.mean(idp)  # the initializer works at unit mean inverse depth
    idn = (idp * (1.0 + idepth_noise * rng.standard_normal(n))).astype(np.float32)
    good = (rng.uniform(size=n) > bad_fraction).astype(np.uint8)
    energy = np.stack([rng.uniform(0, 400, n), rng.uniform(0, 0.2, n)], axis=1).astype(np.float32)
    return dict(u=(xx + 0.1).astype(np.float32), v=(yy + 0.1).astype(np.float32), idepth_new=idn,
                iR=(idp * (1.0 + 0.02 * rng.standard_normal(n))).astype(np.float32), isGood=good, energy=energy,
                outlierTH=np.full(n, 8 * 12.0 * 12.0, dtype=np.float32))


def level_K(K, lvl):
    """fx, fy, cx, cy of pyramid level lvl with CoarseInitializer::makeK's float arithmetic (:963-976)."""
    fx, fy, cx, cy = (np.float32(k) for k in K)
    fxl, fyl = fx, fy
    for _ in range(lvl):
        fxl = np.float32(np.float64(fxl) * 0.5)
        fyl = np.float32(np.float64(fyl) * 0.5)
    if lvl == 0:
        return np.array([fx, fy, cx, cy], dtype=np.float32)
    cxl = np.float32((np.float64(cx) + 0.5) / (1 << lvl) - 0.5)
    cyl = np.float32((np.float64(cy) + 0.5) / (1 << lvl) - 0.5)
    return np.array([fxl, fyl, cxl, cyl], dtype=np.float32)


# ---------------------------------------------------------------- ImmaturePoint tracing
def trace_geometry(K, pose7, aff=(0.0, 0.0)):
    """hostToFrame_KRKi, hostToFrame_Kt, hostToFrame_affine as FullSystem::traceNewCoarse forms them (FullSystem.cpp:717-721)
    for exposure 1 and a host with aff_g2l = 0: float K * float R * K^-1, K * float t, (exp(a), b)."""
    fx, fy, cx, cy = K
    Kf = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float32)
    Rf = quat_to_R(pose7[:4]).astype(np.float32)
    tf = np.asarray(pose7[4:7], dtype=np.float32)
    Kif = np.linalg.inv(Kf.astype(np.float64)).astype(np.float32)
    KRKi = ((Kf @ Rf).astype(np.float32) @ Kif).astype(np.float32)
    Kt = (Kf @ tf).astype(np.float32)
    return KRKi, Kt, np.array([np.exp(aff[0]), aff[1]], dtype=np.float32)


def immature_candidates(scene: Scene, step=7, border=8, seed=DEFAULT_SEED):
    """Integer pixel positions on a jittered grid (stand-in for the selector's picks) + their ground-truth inverse depth."""
    rng = np.random.default_rng(seed)
    xs = np.arange(border, scene.w - border, step)
    ys = np.arange(border, scene.h - border, step)
    xx, yy = np.meshgrid(xs, ys)
    xx = (xx + rng.integers(0, max(step - 1, 1), xx.shape)).ravel().clip(border, scene.w - border - 1)
    yy = (yy + rng.integers(0, max(step - 1, 1), yy.shape)).ravel().clip(border, scene.h - border - 1)
    return xx.astype(np.float32), yy.astype(np.float32), scene.idepth(xx, yy).astype(np.float32)
