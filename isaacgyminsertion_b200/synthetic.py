"""Synthetic stand-ins for the closed IsaacGym dependency.

IsaacGym physics and its Vulkan camera sensors cannot run here, so the bench and
the parity tests drive the observation path with:
  * `SyntheticGym`           the three gym calls CameraPointCloud.__init__ makes
                             (pcl_utils.py:122-135): view / projection matrix, env origin;
  * `external_camera_frames` analytic depth + segmentation frames of a table plane,
                             a socket box (seg 3), a plug box (seg 2) and an arm blob (seg 1)
                             seen from the task's real camera pose
                             (FactoryEnvInsertionTactile.yaml:39-40), per-env hfov 70 +- 5 deg
                             (factory_env_insertion.py:516), depth negative, -inf on ray miss;
  * `tactile_poses`          fingertip / plug poses with a stated contact mix.

Everything is seeded and pure numpy; nothing here reads /root/reference.
"""
import math
from types import SimpleNamespace

import numpy as np

REAL_CAM_POS = np.array([0.73114316, -0.01966786, 0.1629284])
REAL_CAM_ORI_XYZW = np.array([0.60720, 0.6214361, -0.3433028, -0.3567319])  # optical frame
CAM_W, CAM_H = 96, 54
ENV_SPACING = 0.5  # FactoryBaseTactile.yaml:39

SEG_TABLE, SEG_KUKA, SEG_PLUG, SEG_SOCKET = 0, 1, 2, 3  # factory_env_insertion.py:814-848


def quat_to_matrix(q):
    """xyzw quaternion(s) -> rotation matrix, normalising like scipy R.from_quat."""
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def matrix_to_quat(R):
    """rotation matrix -> xyzw quaternion (w >= 0)."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = [0.0, 0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    q = np.array(q)
    return q if q[3] >= 0 else -q


def env_origins(n_envs, spacing=ENV_SPACING):
    """gym.create_env grid: int(sqrt(N)) envs per row, cell size 2*spacing."""
    per_row = max(int(math.sqrt(n_envs)), 1)
    i = np.arange(n_envs)
    o = np.zeros((n_envs, 3))
    o[:, 0] = (i % per_row) * 2 * spacing
    o[:, 1] = (i // per_row) * 2 * spacing
    return o


class SyntheticGym:
    """Duck-typed subset of isaacgym's gym used by CameraPointCloud.__init__."""

    def __init__(self, n_envs, seed=0, width=CAM_W, height=CAM_H, global_env_offset=0, total_envs=None):
        rng = np.random.default_rng(seed)
        total = total_envs if total_envs is not None else n_envs
        hfov_all = 70 + rng.integers(-5, 6, size=total)
        origins_all = env_origins(total)
        sl = slice(global_env_offset, global_env_offset + n_envs)
        self.n_envs = n_envs
        self.global_env_offset = global_env_offset
        self.width, self.height = width, height
        self.hfov = hfov_all[sl].astype(np.float64)
        self.origins = origins_all[sl]
        R_opt = quat_to_matrix(REAL_CAM_ORI_XYZW)
        self.R_gl = R_opt @ np.diag([1.0, -1.0, -1.0])  # GL camera: x right, y up, z back
        self.envs = list(range(n_envs))
        self.camera_handles = list(range(n_envs))
        self.camera_props = [SimpleNamespace(width=width, height=height, horizontal_fov=float(h))
                             for h in self.hfov]
        self._view = []
        self._proj = []
        for e in range(n_envs):
            M = np.eye(4)
            M[:3, :3] = self.R_gl
            M[:3, 3] = REAL_CAM_POS + self.origins[e]
            V = np.linalg.inv(M).T  # row-vector convention: p_cam = p_world @ V
            self._view.append(V.astype(np.float32))
            p00 = 1.0 / math.tan(math.radians(self.hfov[e]) / 2)
            P = np.zeros((4, 4), dtype=np.float32)
            P[0, 0] = p00
            P[1, 1] = p00 * width / height
            P[2, 2] = -1.0
            P[2, 3] = -1.0
            P[3, 2] = -0.02
            self._proj.append(P)

    def get_camera_view_matrix(self, sim, env, handle):
        return self._view[env]

    def get_camera_proj_matrix(self, sim, env, handle):
        return self._proj[env]

    def get_env_origin(self, env):
        o = self.origins[env]
        return SimpleNamespace(x=float(o[0]), y=float(o[1]), z=float(o[2]))


def _ray_box(o, d, center, R, half):
    """Ray/OBB slab test. o (3,), d (P,3) world; returns t (P,) (inf on miss)."""
    ol = R.T @ (o - center)
    dl = d @ R
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dl
        t1 = (-half - ol) * inv
        t2 = (half - ol) * inv
    tmin = np.nanmax(np.minimum(t1, t2), axis=1)
    tmax = np.nanmin(np.maximum(t1, t2), axis=1)
    hit = (tmax >= np.maximum(tmin, 0.0))
    return np.where(hit, np.maximum(tmin, 0.0), np.inf)


def _ray_sphere(o, d, c, r):
    oc = o - c
    b = d @ oc
    cc = oc @ oc - r * r
    dd = np.einsum("ij,ij->i", d, d)
    disc = b * b - dd * cc
    t = (-b - np.sqrt(np.maximum(disc, 0.0))) / dd
    return np.where((disc >= 0) & (t > 0), t, np.inf)


def external_camera_frames(gym, plug_pos, plug_quat, socket_pos, seed=0, miss_fraction=0.02):
    """Analytic depth/seg for every env of `gym`.

    plug_pos (N,3), plug_quat (N,4 xyzw), socket_pos (N,3) are env-local.
    Returns depth (N,H,W) f32 (negative metric z, -inf on miss) and seg (N,H,W) i32.
    """
    N, H, W = gym.n_envs, gym.height, gym.width
    g0 = getattr(gym, "global_env_offset", 0)
    depth = np.empty((N, H, W), dtype=np.float32)
    seg = np.empty((N, H, W), dtype=np.int32)
    vv, uu = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    for e in range(N):
        P = gym._proj[e]
        fu = W * P[0, 0] / 2
        fv = H * P[1, 1] / 2
        dcam = np.stack([(uu - W / 2) / fu, -(vv - H / 2) / fv, -np.ones_like(uu, dtype=np.float64)], -1)
        d = dcam.reshape(-1, 3) @ gym.R_gl.T  # world direction, |z_cam| = 1 per unit t
        o = REAL_CAM_POS.copy()  # env-local
        t_best = np.full(H * W, np.inf)
        s_best = np.zeros(H * W, dtype=np.int32)
        # table plane z = 0 (seg 0)
        with np.errstate(divide="ignore"):
            t = -o[2] / d[:, 2]
        t = np.where((d[:, 2] < 0) & (t > 0), t, np.inf)
        t_best, s_best = t, np.where(np.isfinite(t), SEG_TABLE, 0)
        # socket: 5x5x3 cm box sitting on the socket position
        c = socket_pos[e] + np.array([0, 0, 0.015])
        t = _ray_box(o, d, c, np.eye(3), np.array([0.025, 0.025, 0.015]))
        upd = t < t_best
        t_best = np.where(upd, t, t_best)
        s_best = np.where(upd, SEG_SOCKET, s_best)
        # plug: 3 x 3 x 7.62 cm box, origin at its base
        Rp = quat_to_matrix(plug_quat[e])
        c = plug_pos[e] + Rp @ np.array([0, 0, 0.0381])
        t = _ray_box(o, d, c, Rp, np.array([0.015, 0.015, 0.0381]))
        upd = t < t_best
        t_best = np.where(upd, t, t_best)
        s_best = np.where(upd, SEG_PLUG, s_best)
        # arm / hand blob above the plug
        c = plug_pos[e] + Rp @ np.array([0, 0, 0.0762 + 0.06])
        t = _ray_sphere(o, d, c, 0.05)
        upd = t < t_best
        t_best = np.where(upd, t, t_best)
        s_best = np.where(upd, SEG_KUKA, s_best)
        # a few dropped returns (ray misses -> -inf), also on object pixels
        # (drawn per GLOBAL env id, so an env's frame does not depend on how the envs are sharded)
        drop = np.random.default_rng([seed + 7919, g0 + e]).random(H * W) < miss_fraction
        t_best = np.where(drop, np.inf, t_best)
        depth[e] = (-t_best).astype(np.float32).reshape(H, W)
        seg[e] = s_best.reshape(H, W)
    return depth, seg


def scene_poses(n_envs, seed=0, assets=None, global_env_offset=0):
    """Plug / socket poses per env.  With the packed asset file they cycle through the
    reference's grasp tables (initial_grasp_data/<sub>_noise.npz rows);
    otherwise they are drawn from the ranges those tables span."""
    rng = np.random.default_rng(seed + 104729)
    gid = np.arange(global_env_offset, global_env_offset + n_envs)
    plug_pos = np.empty((n_envs, 3))
    plug_quat = np.empty((n_envs, 4))
    socket_pos = np.empty((n_envs, 3))
    if assets is not None:
        for k, g in enumerate(gid):
            sub = g % 7
            row = (g // 7) % len(assets[f"grasp_{sub}_plug_pos"])
            plug_pos[k] = assets[f"grasp_{sub}_plug_pos"][row]
            plug_quat[k] = assets[f"grasp_{sub}_plug_quat"][row]
            socket_pos[k] = assets[f"grasp_{sub}_socket_pos"][row]
    else:
        all_rng = np.random.default_rng(seed + 104729)
        tot = global_env_offset + n_envs
        sp = np.stack([all_rng.uniform(0.45, 0.55, tot), all_rng.uniform(-0.05, 0.05, tot),
                       all_rng.uniform(0.001, 0.051, tot)], 1)
        pp = sp + np.stack([all_rng.uniform(-0.01, 0.01, tot), all_rng.uniform(-0.01, 0.01, tot),
                            all_rng.uniform(0.03, 0.06, tot)], 1)
        qv = all_rng.uniform(-0.18, 0.18, (tot, 3))
        pq = np.concatenate([qv, np.ones((tot, 1))], 1)
        pq /= np.linalg.norm(pq, axis=1, keepdims=True)
        plug_pos, plug_quat, socket_pos = pp[gid], pq[gid], sp[gid]
    del rng
    return plug_pos, plug_quat, socket_pos


# ---------------------------------------------------------------------------
# tactile poses
# ---------------------------------------------------------------------------
GEL_OUTER_R = 0.0140     # outer wall radius of the allsight dome (x,y scaled mesh)
SENSOR_MID_X = 0.020     # sensor mid-length along its own +x axis
PLUG_LENGTH = 0.0762


def _support_radius(verts, h, phi, band=0.004):
    """Plug cross-section support distance in direction phi near height h (plug frame)."""
    sel = np.abs(verts[:, 2] - h) < band
    v = verts[sel] if sel.any() else verts
    return float((v[:, 0] * math.cos(phi) + v[:, 1] * math.sin(phi)).max())


def _axis_angle(axis, ang):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def tactile_poses(n_envs, assets, seed=0, global_env_offset=0, contact_mix=(0.5, 0.25, 0.25)):
    """Seeded fingertip / plug poses for N envs x 3 fingertip sensors.

    Per env (global id g): plug mesh g % 7, plug pose from the grasp table; per finger k a
    sensor whose axis (+x, camera looks along it) is anti-parallel to the plug axis, tilted
    U(0,40) deg towards the plug, rolled U(0,2pi), its mid-length on a ring at 0.9*plug length,
    azimuth k*120 + U(-10,10) deg, with the plug surface penetrating the gel's outer wall by
    delta: `contact_mix` = fractions of (deep U(3,6) mm, grazing U(0,3) mm, none U(-3,0) mm).
    Returns dict of f32 arrays: finger_pos (N,3,3), finger_quat (N,3,4 xyzw), plug_pos (N,3),
    plug_quat (N,4), mesh_id (N) i32, bg_id (N,3) i32 in 12..19, delta (N,3).
    """
    out = dict(finger_pos=np.empty((n_envs, 3, 3), np.float32), finger_quat=np.empty((n_envs, 3, 4), np.float32),
               plug_pos=np.empty((n_envs, 3), np.float32), plug_quat=np.empty((n_envs, 4), np.float32),
               mesh_id=np.empty(n_envs, np.int32), bg_id=np.empty((n_envs, 3), np.int32),
               delta=np.empty((n_envs, 3), np.float32))
    n_pegs = len(assets["peg_names"])
    verts = [np.asarray(assets[f"peg_{i}_v"], dtype=np.float64) for i in range(n_pegs)]
    h = 0.9 * PLUG_LENGTH
    for k in range(n_envs):
        g = global_env_offset + k
        rng = np.random.default_rng([seed, g])          # counter-based: independent of the sharding
        mid = g % n_pegs
        row = (g // n_pegs) % len(assets[f"grasp_{mid}_plug_pos"])
        ppos = np.asarray(assets[f"grasp_{mid}_plug_pos"][row], dtype=np.float64)
        pquat = np.asarray(assets[f"grasp_{mid}_plug_quat"][row], dtype=np.float64)
        Rp = quat_to_matrix(pquat)
        out["plug_pos"][k] = ppos
        out["plug_quat"][k] = pquat / np.linalg.norm(pquat)
        out["mesh_id"][k] = mid
        for f in range(3):
            phi = math.radians(120.0 * f + rng.uniform(-10, 10))
            u = rng.random()
            if u < contact_mix[0]:
                delta = rng.uniform(0.003, 0.006)
            elif u < contact_mix[0] + contact_mix[1]:
                delta = rng.uniform(0.0, 0.003)
            else:
                delta = rng.uniform(-0.003, 0.0)
            tilt = math.radians(rng.uniform(0, 40))
            roll = rng.uniform(0, 2 * math.pi)
            radial = np.array([math.cos(phi), math.sin(phi), 0.0])
            tang = np.array([-math.sin(phi), math.cos(phi), 0.0])
            rho = _support_radius(verts[mid], h, phi) + GEL_OUTER_R - delta
            p_mid = radial * rho + np.array([0, 0, h])
            # sensor axes in the plug frame: x_s = -z tilted towards the plug axis about the tangent
            x_s = _axis_angle(tang, -tilt) @ np.array([0.0, 0.0, -1.0])
            if x_s @ radial > 0:                          # make sure the tip leans towards the plug
                x_s = _axis_angle(tang, tilt) @ np.array([0.0, 0.0, -1.0])
            y0 = tang
            z0 = np.cross(x_s, y0)
            Rs = np.stack([x_s, y0, z0], axis=1) @ _axis_angle([1, 0, 0], roll)
            origin = p_mid - SENSOR_MID_X * x_s
            Rw = Rp @ Rs
            out["finger_pos"][k, f] = ppos + Rp @ origin
            out["finger_quat"][k, f] = matrix_to_quat(Rw)
            out["bg_id"][k, f] = 12 + int(rng.integers(0, 8))
            out["delta"][k, f] = delta
    return out
