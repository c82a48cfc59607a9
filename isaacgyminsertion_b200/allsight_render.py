"""Batched allsight tactile renderer + drop-in `AllSightRenderer` handles.

Reference surface kept (isaacgyminsertion/allsight/experiments/allsight_render.py:50-219):
`AllSightRenderer.{update_pose_given_sim_pose, render, remove_bg, get_background, bg_img,
bg_depth, mask, renderer.depth0}`; the reference builds one object (one GL context, one copy
of the 231k-triangle gel) per env x fingertip (factory_env_insertion.py:1047-1053) and renders
them one by one (factory_task_insertion.py:521-531).  Here ONE `BatchedAllSight` owns every
sensor of the rank's env slice and renders all frames with a few kernel launches; the
per-sensor handles it hands out keep the reference's method names and return types.

No CPU fallback: the device must be CUDA and libigi_b200.so must be built.
"""
import ctypes
import math
import os

import numpy as np
import torch
import yaml
from scipy import ndimage
from scipy.spatial.transform import Rotation as R

from . import _lib
from . import assets as _assets

_c = ctypes
W = H = 224
OBS_W, OBS_H = 64, 32
OBS_LEN = OBS_W * OBS_H


# --- ctypes mirrors of include/igi_b200.h ----------------------------------------------------
class IgiSensorParams(_c.Structure):
    _fields_ = [
        ("width", _c.c_int32), ("height", _c.c_int32), ("znear", _c.c_float),
        ("dxp_first", _c.c_float), ("dxp_last", _c.c_float), ("dyp_first", _c.c_float), ("dyp_last", _c.c_float),
        ("n_lights", _c.c_int32),
        ("light_pos", _c.c_void_p), ("light_dir", _c.c_void_p), ("light_col", _c.c_void_p),
        ("light_int", _c.c_void_p), ("light_las", _c.c_void_p), ("light_lao", _c.c_void_p),
        ("inverse_square", _c.c_int32),
        ("base_color", _c.c_float * 3), ("metallic", _c.c_float), ("roughness", _c.c_float),
        ("cam_R", _c.c_double * 9), ("cam_p", _c.c_double * 3),
        ("max_force", _c.c_double), ("max_deformation", _c.c_double),
        ("calib_scale", _c.c_float), ("clip_lo", _c.c_float), ("clip_hi", _c.c_float),
        ("blur_ksize", _c.c_int32), ("gauss", _c.c_float * 7),
        ("grid_org", _c.c_float * 3), ("grid_h", _c.c_float), ("grid_slack", _c.c_float),
        ("grid_n", _c.c_int32 * 3), ("depth0_max", _c.c_float),
        ("hiz_levels", _c.c_int32), ("hiz_off", _c.c_int32 * 10), ("hiz_w", _c.c_int32 * 10),
    ]


class IgiTactileMeshes(_c.Structure):
    _fields_ = [("verts", _c.c_void_p), ("vnorm", _c.c_void_p), ("faces", _c.c_void_p),
                ("face_orig", _c.c_void_p), ("meshes", _c.c_void_p), ("clusters", _c.c_void_p)]


class IgiTactileStatic(_c.Structure):
    _fields_ = [("depth0", _c.c_void_p), ("bg_sim", _c.c_void_p), ("bg_real", _c.c_void_p),
                ("obs_empty", _c.c_void_p), ("grid", _c.c_void_p), ("hiz", _c.c_void_p),
                ("dxp", _c.c_void_p), ("dyp", _c.c_void_p)]


class IgiTactileFrames(_c.Structure):
    _fields_ = [("n_envs", _c.c_int32), ("sensors_per_env", _c.c_int32),
                ("finger_pos", _c.c_void_p), ("finger_quat", _c.c_void_p),
                ("plug_pos", _c.c_void_p), ("plug_quat", _c.c_void_p),
                ("force", _c.c_void_p), ("force_const", _c.c_float),
                ("update", _c.c_void_p), ("update2", _c.c_void_p),
                ("finger_pos_n", _c.c_void_p * 8), ("finger_quat_n", _c.c_void_p * 8),
                ("finger_pos_stride", _c.c_int64), ("finger_quat_stride", _c.c_int64),
                ("mesh_id", _c.c_void_p), ("bg_id", _c.c_void_p),
                ("stage_mask", _c.c_int32), ("region_budget", _c.c_int32), ("fill_split", _c.c_int32)]


class IgiTactileScratch(_c.Structure):
    _fields_ = [("M", _c.c_void_p), ("setups", _c.c_void_p), ("normals", _c.c_void_p), ("counts", _c.c_void_p), ("bbox", _c.c_void_p),
                ("worklist", _c.c_void_p), ("counters", _c.c_void_p), ("kmax", _c.c_int32)]


class IgiTactileOut(_c.Structure):
    _fields_ = [("color", _c.c_void_p), ("gel_depth", _c.c_void_p), ("obs", _c.c_void_p),
                ("obs_env_stride", _c.c_int64), ("obs_sensor_stride", _c.c_int64)]


def euler2matrix(angles=(0, 0, 0), translation=(0, 0, 0), xyz="xyz", degrees=False):
    pose = np.eye(4)
    pose[:3, 3] = translation
    pose[:3, :3] = R.from_euler(xyz, angles, degrees=degrees).as_matrix()
    return pose


def circle_mask(size=(224, 224), border=0):
    """Same mask as allsight_render.py:28-40 (cv2.circle filled, radius min(size)//2),
    which equals (x-cx)^2+(y-cy)^2 <= r^2 on integer pixel coordinates."""
    yy, xx = np.mgrid[0:size[1], 0:size[0]]
    r = min(size[0], size[1]) // 2 - border
    m = (((xx - size[0] // 2) ** 2 + (yy - size[1] // 2) ** 2) <= r * r).astype(np.float32)
    return np.stack([m, m, m], axis=2)


class SensorConfig:
    """Constants of the sensor yaml in the camera frame (tacto/renderer.py:291-325,
    tacto_allsight_wrapper/allsight_wrapper.py:100-174)."""

    def __init__(self, yml=_assets.SENSOR_YML, falloff=None):
        conf = yaml.safe_load(open(yml))["sensor"]
        cam = conf["camera"][0]
        self.cam_zero = euler2matrix(angles=np.deg2rad(cam["orientation"]), translation=cam["position"])
        self.znear = float(cam["znear"])
        t = math.tan(math.radians(cam["yfov"]) / 2.0)
        px = np.arange(W, dtype=np.float64)
        self.dxp = (((px + 0.5) / W * 2.0 - 1.0) * t).astype(np.float32)
        self.dyp = ((1.0 - (px + 0.5) / H * 2.0) * t).astype(np.float32)
        lg = conf["lights"]
        Rc, pc = self.cam_zero[:3, :3], self.cam_zero[:3, 3]
        origin = np.array(lg["origin"], dtype=np.float64)
        inner, outer = np.pi * lg["spot_angles"]["inner"], np.pi * lg["spot_angles"]["outer"]
        pos, direc = [], []
        for i, th in enumerate(lg["xrtheta"]["thetas"]):
            theta = np.pi / 180 * th
            xyz = np.array([lg["xrtheta"]["xs"][i], lg["xrtheta"]["rs"][i] * np.cos(theta),
                            lg["xrtheta"]["rs"][i] * np.sin(theta)])
            pose = euler2matrix(xyz="yzx", angles=[-np.pi / 16, 0, np.pi / 180 * (th - 90)], translation=xyz + origin)
            pos.append(Rc.T @ (pose[:3, 3] - pc))
            direc.append(Rc.T @ (-pose[:3, 2]))
        n = len(pos)
        las = 1.0 / max(0.001, np.cos(inner) - np.cos(outer))
        self.light_pos = np.ascontiguousarray(pos, dtype=np.float32)
        self.light_dir = np.ascontiguousarray(direc, dtype=np.float32)
        self.light_col = np.ascontiguousarray(lg["colors"], dtype=np.float32)
        self.light_int = np.ascontiguousarray(lg["intensities"], dtype=np.float32)
        self.light_las = np.full(n, las, dtype=np.float32)
        self.light_lao = np.full(n, -np.cos(outer) * las, dtype=np.float32)
        # pyrender attenuates punctual lights by 1/d^2 (KHR_lights_punctual); `none` is the documented
        # deviation (DESIGN.md "light model").  `falloff` overrides the yaml key.
        self.falloff = falloff if falloff is not None else lg.get("falloff", "inverse_square")
        if self.falloff not in ("inverse_square", "none"):
            raise ValueError(f"lights.falloff must be inverse_square or none, got {self.falloff!r}")
        self.inverse_square = 1 if self.falloff == "inverse_square" else 0
        self.max_force = float(conf["force"]["range_force"][1])
        self.max_deformation = float(conf["force"]["max_deformation"])
        cal = conf["bg_calibration"]
        self.calib_scale = float(cal["scale_factor"])
        self.clip = (float(cal["clip"][0]), float(cal["clip"][1]))
        k, sig = int(cal["blur"]["k_size"]), float(cal["blur"]["sigma"])
        g = np.exp(-((np.arange(k) - (k - 1) / 2.0) ** 2) / (2.0 * sig * sig))
        self.gauss = (g / g.sum())
        self.blur_ksize = k
        m = conf["material"]
        self.base_color = [float(v) for v in m["base_color"]]
        self.metallic, self.roughness = float(m["metallic"]), float(m["roughness"])


def _morton3(q):
    def part(v):
        v = v.astype(np.uint64) & 0x3ff
        v = (v | (v << 16)) & 0x30000ff
        v = (v | (v << 8)) & 0x300f00f
        v = (v | (v << 4)) & 0x30c30c3
        v = (v | (v << 2)) & 0x9249249
        return v
    return part(q[:, 0]) | (part(q[:, 1]) << 1) | (part(q[:, 2]) << 2)


class MeshTable:
    """All plug meshes concatenated; faces Morton-sorted and cut into clusters of
    `cluster_size` with a bounding sphere each (the geometry kernel culls clusters first)."""

    def __init__(self, meshes, cluster_size=64):
        verts, vnorm, faces, orig, info, clusters = [], [], [], [], [], []
        voff = foff = 0
        for (V, VN, F) in meshes:
            V = np.asarray(V, dtype=np.float32)
            F = np.asarray(F, dtype=np.int32)
            cen = V[F].mean(axis=1)
            lo, hi = cen.min(0), cen.max(0)
            q = ((cen - lo) / np.maximum(hi - lo, 1e-12) * 1023.0).astype(np.int64)
            order = np.argsort(_morton3(q), kind="stable")
            Fs = F[order]
            cl_off = len(clusters)
            for s in range(0, len(Fs), cluster_size):
                fc = Fs[s:s + cluster_size]
                P = V[fc].reshape(-1, 3).astype(np.float64)
                c = 0.5 * (P.min(0) + P.max(0))
                r = float(np.sqrt(((P - c) ** 2).sum(1).max())) * 1.0001 + 1e-7
                clusters.append((c[0], c[1], c[2], r, foff + s, len(fc)))
            info.append((foff, len(Fs), cl_off, len(clusters) - cl_off))
            verts.append(V)
            vnorm.append(np.asarray(VN, dtype=np.float32))
            faces.append(Fs + voff)
            orig.append(order.astype(np.int32))
            voff += len(V)
            foff += len(Fs)
        self.verts = np.concatenate(verts)
        self.vnorm = np.concatenate(vnorm)
        self.faces = np.concatenate(faces).astype(np.int32)
        self.face_orig = np.concatenate(orig).astype(np.int32)
        self.info = np.asarray(info, dtype=np.int32)
        cl = np.zeros(len(clusters), dtype=[("c", np.float32, 4), ("i", np.int32, 4)])
        for k, (cx, cy, cz, r, first, cnt) in enumerate(clusters):
            cl["c"][k] = (cx, cy, cz, r)
            cl["i"][k] = (first, cnt, 0, 0)
        self.clusters = cl
        self.max_clusters = int(self.info[:, 3].max())


def depth_max_pyramid(depth0):
    """Max-pyramid of depth0 for hierarchical-Z culling: (flat f32, offsets, widths)."""
    cur = np.where(depth0 > 0, depth0, np.inf).astype(np.float32)
    levels, offs, widths, off = [], [], [], 0
    while True:
        levels.append(cur.ravel())
        offs.append(off)
        widths.append(cur.shape[1])
        off += cur.size
        if cur.shape[0] == 1 and cur.shape[1] == 1:
            break
        hh, ww = (cur.shape[0] + 1) // 2, (cur.shape[1] + 1) // 2
        pad = np.full((hh * 2, ww * 2), -np.inf, dtype=np.float32)
        pad[:cur.shape[0], :cur.shape[1]] = cur
        cur = pad.reshape(hh, 2, ww, 2).max(axis=(1, 3))
    return np.concatenate(levels).astype(np.float32), offs, widths


def gel_interior_grid(depth0, dxp, dyp, h=0.00025):
    """Conservative distance grid to the visible gel interior {points nearer than depth0
    along their pixel ray}, camera frame (x right, y up, depth forward).  Returns
    (grid f32 (nz,ny,nx) metres, origin (3,), h, slack)."""
    d0 = depth0.astype(np.float32)
    d0 = np.where(d0 > 0, d0, d0.max())
    ss = np.linspace(0.0, 1.0, 128, dtype=np.float32)   # <= 0.2 mm between samples along a ray
    X = (dxp[None, :, None].astype(np.float32) * d0[:, :, None] * ss)
    Y = (dyp[:, None, None].astype(np.float32) * d0[:, :, None] * ss)
    Z = d0[:, :, None] * ss
    lo = np.array([X.min(), Y.min(), 0.0]) - 2 * h
    hi = np.array([X.max(), Y.max(), Z.max()]) + 2 * h
    n = np.ceil((hi - lo) / h).astype(int)
    occ = np.zeros((n[2], n[1], n[0]), dtype=bool)
    ix = ((X - lo[0]) / h).astype(int).ravel()
    iy = ((Y - lo[1]) / h).astype(int).ravel()
    iz = ((Z - lo[2]) / h).astype(int).ravel()
    occ[iz, iy, ix] = True
    occ = ndimage.binary_dilation(occ, structure=np.ones((3, 3, 3), dtype=bool))
    dist = ndimage.distance_transform_edt(~occ) * h
    slack = math.sqrt(3.0) * h * 1.01
    return dist.astype(np.float32), lo.astype(np.float32), float(h), float(slack)


class BatchedAllSight:
    """Every allsight sensor of an env slice, rendered together on one CUDA device."""

    def __init__(self, num_envs, mesh_ids, bg_ids=None, device="cuda", sensors_per_env=3, meshes=None,
                 sensor_yml=_assets.SENSOR_YML, assets_path=_assets.ASSETS_NPZ, kmax=1024, seed=None,
                 falloff=None, on_overflow="grow"):
        """kmax: triangle-list capacity per frame (scratch = F * kmax * 112 B).  A frame that produces more
        candidate triangles sets a device flag; `on_overflow` says what the next `render()` call does when it
        sees the flag of an earlier step (checked without a host sync, see `poll_overflow`): "grow" (default)
        re-allocates the scratch from the measured high-water mark and warns, "raise" raises."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedAllSight needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.N, self.S = int(num_envs), int(sensors_per_env)
        self.F = self.N * self.S
        self.kmax = int(kmax)
        if on_overflow not in ("grow", "raise"):
            raise ValueError("on_overflow must be 'grow' or 'raise'")
        self.on_overflow = on_overflow
        self.capturing = False      # True while a CUDA graph captures render(): no overflow telemetry inside the graph
        self.region_budget = 0      # test hook (IgiTactileFrames.region_budget)
        self.fill_split = int(os.environ.get("IGI_FILL_SPLIT", "0"))   # tuning hook (IgiTactileFrames.fill_split)
        self.cfg = SensorConfig(sensor_yml, falloff=falloff)
        packed = _assets.load_packed(assets_path)
        if meshes is None:
            meshes = [(packed[f"peg_{i}_v"], packed[f"peg_{i}_vn"], packed[f"peg_{i}_f"])
                      for i in range(len(packed["peg_names"]))]
        self.mesh_table = MeshTable(meshes)
        dev = self.device
        mt = self.mesh_table
        self._verts = torch.from_numpy(mt.verts).to(dev)
        self._vnorm = torch.from_numpy(mt.vnorm).to(dev)
        self._faces = torch.from_numpy(mt.faces).to(dev)
        self._face_orig = torch.from_numpy(mt.face_orig).to(dev)
        self._minfo = torch.from_numpy(mt.info).to(dev)
        self._clusters = torch.from_numpy(mt.clusters.view(np.uint8).reshape(-1)).to(dev)
        self._gel = torch.from_numpy(np.ascontiguousarray(packed["gel_tris"], dtype=np.float32)).to(dev)
        self.bg_real = torch.from_numpy(np.ascontiguousarray(packed["bg_real"])).to(dev)
        self.mesh_id = torch.as_tensor(np.asarray(mesh_ids, dtype=np.int32)).to(dev)
        if bg_ids is None:
            # allsight_render.py:70-73: random.randint(12, 19) per handle when randomize=True
            rng = np.random.default_rng(seed)
            bg_ids = 12 + rng.integers(0, 8, size=(self.N, self.S))
        self.bg_ids_host = np.asarray(bg_ids, dtype=np.int32).reshape(self.N, self.S)
        self.bg_index = torch.from_numpy(self.bg_ids_host.reshape(-1) - 12).to(dev)

        # K0 with a provisional (empty) grid, then the real grid from depth0
        self.depth0 = torch.empty((H, W), dtype=torch.float32, device=dev)
        self.bg_sim = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
        self._dxp = torch.from_numpy(self.cfg.dxp).to(dev)
        self._dyp = torch.from_numpy(self.cfg.dyp).to(dev)
        self._grid = torch.zeros((1, 1, 1), dtype=torch.float32, device=dev)
        self._hiz_layout = ([0], [1])
        self._sensor = self._sensor_params(np.zeros(3, np.float32), 1.0, 1.0, (1, 1, 1), 1.0)
        zbuf = torch.empty((H * W,), dtype=torch.int64, device=dev)
        with _lib.on_device(dev):
            rc = self.lib.igi_tactile_gel_precompute(_c.byref(self._sensor), _lib.dptr(self._dxp), _lib.dptr(self._dyp),
                                                     _lib.dptr(self._gel), _c.c_int(self._gel.shape[0]), _lib.dptr(zbuf),
                                                     _lib.dptr(self.depth0), _lib.dptr(self.bg_sim),
                                                     _lib.stream_ptr(dev))
        _lib.check(rc, "igi_tactile_gel_precompute")
        d0 = self.depth0.cpu().numpy()
        grid, org, h, slack = gel_interior_grid(d0, self.cfg.dxp, self.cfg.dyp)
        self._grid = torch.from_numpy(grid).to(dev).contiguous()
        hiz, offs, widths = depth_max_pyramid(d0)
        self._hiz = torch.from_numpy(hiz).to(dev)
        self._hiz_layout = (offs, widths)
        self._sensor = self._sensor_params(org, h, slack, grid.shape[::-1], float(d0.max()))
        self.mask = circle_mask((W, H))

        # persistent outputs (tactile_imgs layout of factory_task_insertion.py:310-314)
        self.color = torch.empty((self.N, self.S, H, W, 3), dtype=torch.uint8, device=dev)
        self.gel_depth = torch.empty((self.N, self.S, H, W), dtype=torch.float32, device=dev)
        self.obs = torch.zeros((self.N, self.S, OBS_LEN), dtype=torch.float32, device=dev)
        # scratch
        self._M = torch.empty((self.F, 12), dtype=torch.float32, device=dev)
        self._counts = torch.zeros((self.F,), dtype=torch.int32, device=dev)
        self._bbox = torch.zeros((self.F, 4), dtype=torch.int32, device=dev)
        self._work = torch.zeros((self.F,), dtype=torch.int32, device=dev)
        self._counters = torch.zeros((4,), dtype=torch.int32, device=dev)
        self._alloc_lists()
        # overflow telemetry: counters[2:4] copied to pinned memory after every render (no host sync)
        self._ovf_host = torch.zeros((2,), dtype=torch.int32).pin_memory()
        self._ovf_event = None
        self._last_args = None
        self.high_water = 0
        # obs of a frame without contact (color == bg_real): exact for every background id
        self.obs_empty = torch.empty((OBS_LEN,), dtype=torch.float32, device=dev)
        zero_id = torch.zeros((1,), dtype=torch.int32, device=dev)
        with _lib.on_device(dev):
            rc = self.lib.igi_tactile_obs(_lib.dptr(self.bg_real), _lib.dptr(self.bg_real), _lib.dptr(zero_id), _c.c_int(1),
                                          _lib.dptr(self.obs_empty), _c.c_int64(OBS_LEN), _lib.stream_ptr(dev))
        _lib.check(rc, "igi_tactile_obs")
        self._structs()
        self._handles = None
        self._stage = None

    # ------------------------------------------------------------------------------------
    def _alloc_lists(self):
        dev = self.device
        self._setups = torch.empty((self.F, self.kmax, 16), dtype=torch.int32, device=dev)
        self._normals = torch.empty((self.F, self.kmax, 12), dtype=torch.float32, device=dev)

    def _sensor_params(self, org, h, slack, n_xyz, d0max):
        """IgiSensorParams of this engine (host struct, passed with every call: the library keeps no sensor state)."""
        c = self.cfg
        p = IgiSensorParams()
        p.width, p.height, p.znear = W, H, c.znear
        p.dxp_first, p.dxp_last = float(c.dxp[0]), float(c.dxp[-1])
        p.dyp_first, p.dyp_last = float(c.dyp[0]), float(c.dyp[-1])
        p.n_lights = len(c.light_int)
        p.light_pos, p.light_dir, p.light_col = c.light_pos.ctypes.data, c.light_dir.ctypes.data, c.light_col.ctypes.data
        p.light_int, p.light_las, p.light_lao = c.light_int.ctypes.data, c.light_las.ctypes.data, c.light_lao.ctypes.data
        p.inverse_square = c.inverse_square
        p.base_color = (_c.c_float * 3)(*c.base_color)
        p.metallic, p.roughness = c.metallic, c.roughness
        p.cam_R = (_c.c_double * 9)(*c.cam_zero[:3, :3].reshape(-1))
        p.cam_p = (_c.c_double * 3)(*c.cam_zero[:3, 3])
        p.max_force, p.max_deformation = c.max_force, c.max_deformation
        p.calib_scale, p.clip_lo, p.clip_hi = c.calib_scale, c.clip[0], c.clip[1]
        p.blur_ksize = c.blur_ksize
        p.gauss = (_c.c_float * 7)(*[float(v) for v in c.gauss])
        p.grid_org = (_c.c_float * 3)(*[float(v) for v in org])
        p.grid_h, p.grid_slack = h, slack
        p.grid_n = (_c.c_int32 * 3)(*[int(v) for v in n_xyz])
        p.depth0_max = d0max
        offs, widths = self._hiz_layout
        p.hiz_levels = len(offs)
        p.hiz_off = (_c.c_int32 * 10)(*(list(offs) + [0] * (10 - len(offs))))
        p.hiz_w = (_c.c_int32 * 10)(*(list(widths) + [1] * (10 - len(widths))))
        return p

    def _structs(self):
        m = IgiTactileMeshes()
        m.verts, m.vnorm, m.faces = self._verts.data_ptr(), self._vnorm.data_ptr(), self._faces.data_ptr()
        m.face_orig, m.meshes, m.clusters = self._face_orig.data_ptr(), self._minfo.data_ptr(), self._clusters.data_ptr()
        st = IgiTactileStatic()
        st.depth0, st.bg_sim, st.bg_real = self.depth0.data_ptr(), self.bg_sim.data_ptr(), self.bg_real.data_ptr()
        st.obs_empty, st.grid, st.hiz = self.obs_empty.data_ptr(), self._grid.data_ptr(), self._hiz.data_ptr()
        st.dxp, st.dyp = self._dxp.data_ptr(), self._dyp.data_ptr()
        sc = IgiTactileScratch()
        sc.M, sc.setups, sc.counts = self._M.data_ptr(), self._setups.data_ptr(), self._counts.data_ptr()
        sc.normals = self._normals.data_ptr()
        sc.bbox, sc.worklist, sc.counters, sc.kmax = (self._bbox.data_ptr(), self._work.data_ptr(),
                                                      self._counters.data_ptr(), self.kmax)
        self._m, self._st, self._sc = m, st, sc

    # ------------------------------------------------------------------------------------
    @torch.no_grad()
    def render(self, finger_pos, finger_quat, plug_pos, plug_quat, force=None, update=None, obs_out=None,
               stage_mask=0, _poll=True, update2=None):
        """One batched pass.  finger_pos (N,S,3), finger_quat (N,S,4 xyzw) f32 CUDA tensors - or a list of S
        per-fingertip views (N,3) / (N,4) with unit inner stride and a common env stride, as the reference holds
        them (left/right/middle_finger_pos, factory_task_insertion.py:481-483): no stack copy; plug_pos (N,3),
        plug_quat (N,4); force: None (=70, factory_task_insertion.py:535), scalar, or (N,S) tensor; update /
        update2: None or (N,) bool/uint8 masks, ANDed (update_freq & update_delay, task :523).
        Fills self.color / self.gel_depth / obs (default self.obs, shape (N,S,2048)); frames of
        envs whose update flag is off are left untouched."""
        _poll = _poll and not self.capturing
        if _poll and stage_mask == 0:
            self.poll_overflow()
        obs = self.obs if obs_out is None else obs_out
        if obs.shape != (self.N, self.S, OBS_LEN) or obs.stride(-1) != 1 or obs.dtype != torch.float32 \
                or not obs.is_cuda or obs.data_ptr() % 16 != 0:
            raise RuntimeError("obs_out must be a (N,S,2048) f32 CUDA tensor with unit inner stride")
        fr = IgiTactileFrames()
        fr.n_envs, fr.sensors_per_env = self.N, self.S
        keep = []

        def poses(x, width, name):
            """-> (packed pointer or None, per-sensor pointers, env stride)"""
            if isinstance(x, (list, tuple)):
                views = list(x)
                ok = len(views) == self.S and all(
                    v.is_cuda and v.dtype == torch.float32 and v.shape == (self.N, width) and v.stride(1) == 1
                    and v.stride(0) == views[0].stride(0) and v.stride(0) >= width for v in views)
                if ok:
                    keep.extend(views)
                    return None, [v.data_ptr() for v in views], views[0].stride(0)
                x = torch.stack([v.float() for v in views], dim=1)
            t = x.reshape(self.F, width)
            if not t.is_contiguous() or t.dtype != torch.float32:
                t = t.contiguous().float()
            keep.append(t)
            return _lib.dptr(t, torch.float32, name).value, [], 0
        p, pn, ps = poses(finger_pos, 3, "finger_pos")
        q, qn, qs = poses(finger_quat, 4, "finger_quat")
        fr.finger_pos, fr.finger_quat = p, q
        for k, v in enumerate(pn):
            fr.finger_pos_n[k] = v
        for k, v in enumerate(qn):
            fr.finger_quat_n[k] = v
        fr.finger_pos_stride, fr.finger_quat_stride = ps, qs
        fr.plug_pos = _lib.dptr(plug_pos, torch.float32, "plug_pos").value
        fr.plug_quat = _lib.dptr(plug_quat, torch.float32, "plug_quat").value
        fr.force_const = 70.0
        if force is None:
            fr.force = None
        elif isinstance(force, torch.Tensor):
            ft = force.reshape(self.F).contiguous().float()
            keep.append(ft)
            fr.force = _lib.dptr(ft, torch.float32, "force").value
        else:
            fr.force = None
            fr.force_const = float(force)
        def mask(m, name):
            if m is None:
                return None
            # a bool tensor is one byte (0 / 1) per env: passed as it is
            up = m.view(torch.uint8) if m.dtype == torch.bool and m.is_contiguous() else m.to(torch.uint8).contiguous()
            keep.append(up)
            return _lib.dptr(up, torch.uint8, name).value
        fr.update, fr.update2 = mask(update, "update"), mask(update2, "update2")
        fr.mesh_id, fr.bg_id = self.mesh_id.data_ptr(), self.bg_index.data_ptr()
        fr.stage_mask = int(stage_mask)
        fr.region_budget, fr.fill_split = int(self.region_budget), int(self.fill_split)
        out = IgiTactileOut()
        out.color, out.gel_depth, out.obs = self.color.data_ptr(), self.gel_depth.data_ptr(), obs.data_ptr()
        out.obs_env_stride, out.obs_sensor_stride = obs.stride(0), obs.stride(1)
        with _lib.on_device(self.device):
            rc = self.lib.igi_tactile_render(_c.byref(self._sensor), _c.byref(self._m), _c.byref(self._st), _c.byref(fr),
                                             _c.byref(self._sc), _c.byref(out), _lib.stream_ptr(self.device))
        _lib.check(rc, "igi_tactile_render")
        if stage_mask == 0 and _poll:
            # overflow flag + high-water mark of this step travel to pinned memory behind the kernels (8 bytes,
            # no host sync); the next render() looks at them
            self._ovf_host.copy_(self._counters[2:4], non_blocking=True)
            self._ovf_event = torch.cuda.Event()
            self._ovf_event.record(torch.cuda.current_stream(self.device))
            self._last_args = (finger_pos, finger_quat, plug_pos, plug_quat, force, update, obs_out, update2)
        return obs

    # ------------------------------------------------------------------------------------
    def poll_overflow(self, wait=False):
        """Look at the overflow flag of the steps rendered so far WITHOUT a host sync (the flag of a step
        is visible once its 8-byte copy has landed; `wait=True` blocks for the last step).  On overflow:
        on_overflow == "raise": RuntimeError; "grow": the lists are re-allocated at 1.5 x the measured
        high-water mark (the retry path) and the last step is rendered again into the same outputs, so at
        most the steps consumed between the overflow and this call saw clipped frames (a warning says so).
        Returns True when an overflow was handled."""
        ev = self._ovf_event
        if ev is None or not (wait or ev.query()):
            return False
        if wait:
            ev.synchronize()
        self._ovf_event = None
        flag, mark = int(self._ovf_host[0]), int(self._ovf_host[1])
        self.high_water = max(self.high_water, mark)
        if not flag:
            return False
        self._counters[2:4].zero_()
        if self.on_overflow == "raise" or mark > 4096:
            raise RuntimeError(f"tactile triangle list overflow: a frame produced {mark} candidate triangles, "
                               f"kmax is {self.kmax} (limit 4096)")
        import warnings
        new = min(max(int(mark * 1.5), self.kmax + 1), 4096)
        warnings.warn(f"tactile triangle lists overflowed ({mark} > kmax {self.kmax}): scratch grown to {new} and the "
                      "last step rendered again; earlier steps since the overflow saw clipped frames")
        self.kmax = new
        self._alloc_lists()
        self._structs()
        if self._last_args is not None:
            fp, fq, pp, pq, force, update, obs_out, update2 = self._last_args
            self.render(fp, fq, pp, pq, force=force, update=update, obs_out=obs_out, _poll=False, update2=update2)
        return True

    def check_overflow(self):
        """Host-synchronous check of the last step (tests, debugging): handles an overflow like `poll_overflow`
        and raises if `on_overflow == "raise"`."""
        return self.poll_overflow(wait=True)

    def contact_counts(self):
        """(N,S) surviving-triangle count per frame of the last render (-1 = not updated)."""
        return self._counts.reshape(self.N, self.S)

    # --- reference-style handles ------------------------------------------------------------
    def handles(self):
        """tactile_handles[e][n] as the env builds them (factory_env_insertion.py:1047-1053)."""
        if self._handles is None:
            self._stage = dict(
                fpos=np.zeros((self.N, self.S, 3), np.float32), fquat=np.tile(np.array([0, 0, 0, 1], np.float32),
                                                                              (self.N, self.S, 1)),
                ppos=np.zeros((self.N, 3), np.float32), pquat=np.tile(np.array([0, 0, 0, 1], np.float32), (self.N, 1)),
                force=np.full((self.N, self.S), 20.0, np.float32), dirty=np.zeros(self.N, dtype=bool))
            self._handles = [[AllSightRenderer(engine=self, env=e, sensor=n) for n in range(self.S)]
                             for e in range(self.N)]
        return self._handles

    def _flush(self):
        st = self._stage
        if not st["dirty"].any():
            return
        dev = self.device
        self.render(torch.from_numpy(st["fpos"]).to(dev), torch.from_numpy(st["fquat"]).to(dev),
                    torch.from_numpy(st["ppos"]).to(dev), torch.from_numpy(st["pquat"]).to(dev),
                    force=torch.from_numpy(st["force"]).to(dev), update=torch.from_numpy(st["dirty"]).to(dev))
        st["dirty"][:] = False


class _RendererView:
    """`handle.renderer.depth0` (allsight_render.py:195)."""

    def __init__(self, engine):
        self._engine = engine

    @property
    def depth0(self):
        return [self._engine.depth0.cpu().numpy()]


class AllSightRenderer:
    """Per-sensor handle with the reference's surface (allsight_render.py:50-219).

    Built by `BatchedAllSight.handles()`; `render()` triggers ONE batched launch for every
    handle whose pose changed since the last launch and returns this sensor's
    (color u8 HxWx3, gel_depth f32 HxW) as numpy arrays like the reference.
    Constructing it directly with the reference's arguments (cfg, obj_path, ..., scale)
    creates a private single-sensor engine for that OBJ."""

    def __init__(self, cfg=None, obj_path=None, obj_pose=None, randomize=False, bg_id=None, headless=False,
                 finger_idx=0, scale=1.08, engine=None, env=0, sensor=0, device="cuda"):
        if engine is None:
            if obj_path is None:
                raise ValueError("obj_path is required when no engine is given")
            import random
            bg = random.randint(12, 19) if randomize else 15     # allsight_render.py:70-73
            mesh = _assets.load_peg_from_obj(obj_path, scale)
            engine = BatchedAllSight(1, [0], [[bg]], device=device, sensors_per_env=1, meshes=[mesh])
            engine.handles()
            env = sensor = 0
        self.engine, self.env, self.sensor = engine, env, sensor
        self.finger_idx = finger_idx
        self.render_config = cfg
        self.renderer = _RendererView(engine)
        self.mask = engine.mask
        self.subtract_bg = True
        self.press_depth = 0.001
        self.randomize_light = False

    @property
    def bg_img(self):
        """Calibrated render without object == the real background frame (diff is exactly 0)."""
        return self.engine.bg_real[self.engine.bg_ids_host[self.env, self.sensor] - 12].cpu().numpy()

    @property
    def bg_depth(self):
        return self.engine.depth0.cpu().numpy()

    def get_background(self, frame="gel"):
        return self.bg_depth

    @staticmethod
    def _to_pos_quat(T):
        T = np.asarray(T, dtype=np.float64)
        return T[:3, 3], R.from_matrix(T[:3, :3]).as_quat()

    def update_pose_given_sim_pose(self, cam_pose, object_pose, is_matrix=True):
        if not is_matrix:
            (op, oe), (cp, ce) = object_pose, cam_pose
            object_pose = euler2matrix(angles=oe, translation=op)
            cam_pose = euler2matrix(angles=ce, translation=cp)
        st = self.engine._stage
        p, q = self._to_pos_quat(cam_pose)
        st["fpos"][self.env, self.sensor], st["fquat"][self.env, self.sensor] = p, q
        p, q = self._to_pos_quat(object_pose)
        st["ppos"][self.env], st["pquat"][self.env] = p, q
        st["dirty"][self.env] = True

    def render(self, object_poses=None, normal_forces=None):
        st = self.engine._stage
        normal_forces = 20 if normal_forces is None else normal_forces
        if object_poses is not None:
            p, q = self._to_pos_quat(object_poses)
            st["ppos"][self.env], st["pquat"][self.env] = p, q
        if st["force"][self.env, self.sensor] != np.float32(normal_forces):
            st["force"][self.env, self.sensor] = normal_forces
            st["dirty"][self.env] = True
        self.engine._flush()
        color = self.engine.color[self.env, self.sensor].cpu().numpy()
        gel_depth = self.engine.gel_depth[self.env, self.sensor].cpu().numpy()
        return color, gel_depth

    def remove_bg(self, img1, img2, offset=0.5):
        diff = np.int32(img1) - np.int32(img2)
        return diff / 255.0 + offset

    def updateGUI(self, colors, depths):
        raise RuntimeError("visualisation is not part of the batched observation path")
