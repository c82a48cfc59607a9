"""CPU oracle for the (T) tactile path — TEST INFRASTRUCTURE ONLY.

Restates, per sensor frame and with the reference's serial structure:
  xyzquat_to_tf_numpy                      tasks/factory_tactile/factory_utils.py:351-365
  update_pose_given_sim_pose               allsight/experiments/allsight_render.py:168-172
  AllSightRenderer.render (+matrix2trans)  allsight_render.py:43-47,179-212
  Renderer.render / adjust_with_force      allsight/tacto/renderer.py:560-603,612-648
  pyrender draw                            renderer.py:642   -> oracle/raster.c (PARITY UNPINNED, see its header)
  _calibrate                               tacto_allsight_wrapper/allsight_wrapper.py:57-98  (real cv2)
  remove_bg / mask / flip / crop / resize / gray
                                           allsight_render.py:214-219, factory_task_insertion.py:546-574 (real cv2)

Pinned stages: every cv2 / scipy stage runs the REAL library the reference calls.
Unpinned stage: the GL rasterisation + shading (third-party pyrender, absent here).
"""
import ctypes
import os

import cv2
import numpy as np
import yaml
from scipy.spatial.transform import Rotation as R

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..")
ASSETS = os.path.join(ROOT, "isaacgyminsertion_b200", "data", "assets.npz")
SENSOR_YML = os.path.join(ROOT, "isaacgyminsertion_b200", "data", "sensor_allsight_white.yml")
RASTER_SO = os.path.join(HERE, "liboracle_raster.so")

W = H = 224


class _OracleCam(ctypes.Structure):
    _fields_ = [("W", ctypes.c_int), ("H", ctypes.c_int), ("znear", ctypes.c_float),
                ("dxp", ctypes.c_void_p), ("dyp", ctypes.c_void_p), ("n_lights", ctypes.c_int),
                ("light_pos", ctypes.c_void_p), ("light_dir", ctypes.c_void_p), ("light_col", ctypes.c_void_p),
                ("light_int", ctypes.c_void_p), ("light_las", ctypes.c_void_p), ("light_lao", ctypes.c_void_p),
                ("base", ctypes.c_float * 3), ("metallic", ctypes.c_float), ("roughness", ctypes.c_float),
                ("inverse_square", ctypes.c_int)]


def euler2matrix(angles=(0, 0, 0), translation=(0, 0, 0), xyz="xyz", degrees=False):
    """tacto/renderer.py:38-45."""
    pose = np.eye(4)
    pose[:3, 3] = translation
    pose[:3, :3] = R.from_euler(xyz, angles, degrees=degrees).as_matrix()
    return pose


def xyzquat_to_tf_numpy(position_quat):
    """factory_utils.py:351-365."""
    position_quat = np.atleast_2d(position_quat)
    N = position_quat.shape[0]
    T = np.zeros((N, 4, 4))
    T[:, 0:3, 0:3] = R.from_quat(position_quat[:, 3:]).as_matrix()
    T[:, :3, 3] = position_quat[:, :3]
    T[:, 3, 3] = 1
    return T


def circle_mask(size=(224, 224), border=0):
    """allsight_render.py:28-40."""
    m = np.zeros((size[1], size[0]))
    m_center = (size[0] // 2, size[1] // 2)
    m_radius = min(size[0], size[1]) // 2 - border
    m = cv2.circle(m, m_center, m_radius, 255, -1)
    m /= 255
    m = m.astype(np.float32)
    return np.stack([m, m, m], axis=2)


class SensorModel:
    """Static scene of one allsight sensor (renderer.py:137-163,291-325; wrapper :100-174)."""

    def __init__(self, yml=SENSOR_YML, assets=ASSETS, falloff=None):
        """`falloff` overrides the yaml's `lights.falloff` (inverse_square | none; DESIGN.md "light model")."""
        conf = yaml.safe_load(open(yml))["sensor"]
        self.conf = conf
        cam = conf["camera"][0]
        self.cam_zero = euler2matrix(angles=np.deg2rad(cam["orientation"]), translation=cam["position"])
        self.znear = float(cam["znear"])
        t = np.tan(np.deg2rad(cam["yfov"]) / 2.0)
        px = np.arange(W, dtype=np.float64)
        self.dxp = (((px + 0.5) / W * 2.0 - 1.0) * t * 1.0).astype(np.float32)
        self.dyp = ((1.0 - (px + 0.5) / H * 2.0) * t).astype(np.float32)
        lg = conf["lights"]
        origin = np.array(lg["origin"], dtype=np.float64)
        Rc = self.cam_zero[:3, :3]
        pc = self.cam_zero[:3, 3]
        pos, direc, las, lao = [], [], [], []
        inner = np.pi * lg["spot_angles"]["inner"]
        outer = np.pi * lg["spot_angles"]["outer"]
        for i, th in enumerate(lg["xrtheta"]["thetas"]):
            theta = np.pi / 180 * th
            xyz = np.array([lg["xrtheta"]["xs"][i], lg["xrtheta"]["rs"][i] * np.cos(theta),
                            lg["xrtheta"]["rs"][i] * np.sin(theta)])
            pose = euler2matrix(xyz="yzx", angles=[-np.pi / 16, 0, np.pi / 180 * (th - 90)],
                                translation=xyz + origin)
            pos.append(Rc.T @ (pose[:3, 3] - pc))
            direc.append(Rc.T @ (-pose[:3, 2]))
            s = 1.0 / max(0.001, np.cos(inner) - np.cos(outer))
            las.append(s)
            lao.append(-np.cos(outer) * s)
        self.light_pos = np.ascontiguousarray(pos, dtype=np.float32)
        self.light_dir = np.ascontiguousarray(direc, dtype=np.float32)
        self.light_col = np.ascontiguousarray(lg["colors"], dtype=np.float32)
        self.light_int = np.ascontiguousarray(lg["intensities"], dtype=np.float32)
        self.light_las = np.ascontiguousarray(las, dtype=np.float32)
        self.light_lao = np.ascontiguousarray(lao, dtype=np.float32)
        self.max_force = float(conf["force"]["range_force"][1])
        self.max_deformation = float(conf["force"]["max_deformation"])
        self.calib = conf["bg_calibration"]
        a = np.load(assets)
        self.assets = a
        self.gel_tris = np.ascontiguousarray(a["gel_tris"], dtype=np.float32)
        self.bg_real = a["bg_real"]
        self.pegs = [(np.ascontiguousarray(a[f"peg_{i}_v"]), np.ascontiguousarray(a[f"peg_{i}_vn"]),
                      np.ascontiguousarray(a[f"peg_{i}_f"])) for i in range(len(a["peg_names"]))]
        self.lib = ctypes.CDLL(RASTER_SO)
        self.lib.igi_oracle_render.restype = ctypes.c_int
        c = _OracleCam()
        c.W, c.H, c.znear = W, H, self.znear
        c.dxp, c.dyp = self.dxp.ctypes.data, self.dyp.ctypes.data
        c.n_lights = len(pos)
        c.light_pos, c.light_dir = self.light_pos.ctypes.data, self.light_dir.ctypes.data
        c.light_col, c.light_int = self.light_col.ctypes.data, self.light_int.ctypes.data
        c.light_las, c.light_lao = self.light_las.ctypes.data, self.light_lao.ctypes.data
        m = conf["material"]
        c.base = (ctypes.c_float * 3)(*m["base_color"])
        c.metallic, c.roughness = m["metallic"], m["roughness"]
        self.falloff = falloff if falloff is not None else lg.get("falloff", "inverse_square")
        assert self.falloff in ("inverse_square", "none"), self.falloff
        c.inverse_square = 1 if self.falloff == "inverse_square" else 0
        self._cam = c
        self.mask = circle_mask((W, H))
        # get_background_sim, renderer.py:165-168
        self.bg_sim, self.depth0, _ = self.draw(None, None)

    def draw(self, peg_id, M):
        """The GL draw: (color u8 HxWx3, depth f32 HxW, kind i8 HxW)."""
        color = np.empty((H, W, 3), dtype=np.uint8)
        depth = np.empty((H, W), dtype=np.float32)
        kind = np.empty((H, W), dtype=np.int8)
        if peg_id is None:
            v = vn = f = None
            nf = 0
            Mp = None
        else:
            v, vn, f = self.pegs[peg_id]
            nf = len(f)
            M = np.ascontiguousarray(M, dtype=np.float32)
            Mp = M.ctypes.data
        rc = self.lib.igi_oracle_render(
            ctypes.byref(self._cam), ctypes.c_void_p(self.gel_tris.ctypes.data), ctypes.c_int(len(self.gel_tris)),
            ctypes.c_float(np.float32(self.cam_zero[0, 3])),
            ctypes.c_void_p(v.ctypes.data if v is not None else None),
            ctypes.c_void_p(vn.ctypes.data if vn is not None else None),
            ctypes.c_void_p(f.ctypes.data if f is not None else None), ctypes.c_int(nf),
            ctypes.c_void_p(Mp), ctypes.c_void_p(color.ctypes.data), ctypes.c_void_p(depth.ctypes.data),
            ctypes.c_void_p(kind.ctypes.data))
        assert rc == 0
        return color, depth, kind

    # -- pose chain (f64, scipy — as the reference) --------------------------------
    def object_in_camera(self, finger_tf, object_tf, force):
        """(3,4) f32 object->camera matrix after the force shift.
        update_camera_pose_from_matrix renderer.py:421-440; AllSightRenderer.render
        allsight_render.py:183-191; adjust_with_force renderer.py:587-603."""
        cam_pose = finger_tf.dot(self.cam_zero)
        r = R.from_matrix(object_tf[:3, :3])
        euler = r.as_euler(seq="xyz")                      # matrix2trans, allsight_render.py:43-47
        obj_pos = np.array(object_tf[:3, 3])
        offset = min(self.max_force, force) / self.max_force
        camera_pos = np.array(cam_pose[:3, 3].T)
        direction = camera_pos - obj_pos
        direction = direction / (np.sum(direction ** 2) ** 0.5 + 1e-6)
        obj_pos = obj_pos + offset * self.max_deformation * direction
        pose = euler2matrix(angles=euler, translation=obj_pos)   # update_object_pose renderer.py:445-452
        M = np.linalg.inv(cam_pose) @ pose
        return M[:3, :4].astype(np.float32)

    # -- calibration (real cv2) ------------------------------------------------------
    def calibrate(self, color, bg_real):
        """allsight_wrapper.py:57-98."""
        cfg = self.calib
        diff = (color.astype(np.float64) - self.bg_sim) * cfg["scale_factor"]
        k = cfg["blur"]["k_size"]
        diff = cv2.GaussianBlur(diff, (k, k), cfg["blur"]["sigma"])
        return np.clip((diff[:, :, :3] + bg_real), cfg["clip"][0], cfg["clip"][1]).astype(np.uint8)


class OracleAllSight:
    """One sensor handle: AllSightRenderer (allsight_render.py:50-219)."""

    def __init__(self, model, peg_id, bg_id=15):
        self.model = model
        self.peg_id = peg_id
        self.bg_real = model.bg_real[bg_id - 12]
        self.mask = model.mask
        # calibrated render without object == bg_real (diff is exactly 0)
        self.bg_img = model.calibrate(model.bg_sim, self.bg_real)
        self.bg_depth = model.depth0
        self.finger_tf = np.eye(4)
        self.object_tf = np.eye(4)

    def update_pose_given_sim_pose(self, cam_pose, object_pose):
        self.finger_tf = np.array(cam_pose, dtype=np.float64)
        self.object_tf = np.array(object_pose, dtype=np.float64)

    def render(self, object_poses=None, normal_forces=None, return_raw=False):
        normal_forces = 20 if normal_forces is None else normal_forces
        obj = self.object_tf if object_poses is None else np.asarray(object_poses, dtype=np.float64)
        M = self.model.object_in_camera(self.finger_tf, obj, normal_forces)
        raw, depth, kind = self.model.draw(self.peg_id, M)
        color = self.model.calibrate(raw, self.bg_real)
        gel_depth = self.model.depth0 - depth               # allsight_render.py:193-197
        if return_raw:
            return color, gel_depth, raw, kind, M
        return color, gel_depth

    @staticmethod
    def remove_bg(img1, img2, offset=0.5):
        """allsight_render.py:214-219."""
        diff = np.int32(img1) - np.int32(img2)
        return diff / 255.0 + offset


def tactile_obs(color, bg_img, mask, enc_w=32, enc_h=64):
    """factory_task_insertion.py:546-574 with diff=True, crop_roi=True, num_channels=1 -> (2048,) f32."""
    img = OracleAllSight.remove_bg(color, bg_img)
    img *= mask
    img = np.flipud(img).copy()
    w = img.shape[0]
    img = img[:w // 2, :, :]
    if img.shape[:2] != (enc_w, enc_h):
        img = cv2.resize(img, (enc_h, enc_w), interpolation=cv2.INTER_AREA)
    gray = cv2.cvtColor(img.astype('float32'), cv2.COLOR_BGR2GRAY)
    return gray.flatten()


def render_tactile_serial(model, handles, finger_poses7, object_poses7, forces):
    """_render_tactile (factory_task_insertion.py:515-583): serial loop over envs x 3 sensors.
    handles[e][n]; finger_poses7 (N,3,7) f32; object_poses7 (N,7) f32 -> (N,3,2048) f32."""
    N = len(handles)
    out = np.zeros((N, 3, 2048), dtype=np.float32)
    obj_tf = xyzquat_to_tf_numpy(np.asarray(object_poses7))
    for e in range(N):
        for n in range(3):
            ftf = xyzquat_to_tf_numpy(np.asarray(finger_poses7[e, n]))[0]
            h = handles[e][n]
            h.update_pose_given_sim_pose(ftf, obj_tf[e])
            color, _ = h.render(obj_tf[e], forces[e][n] if np.ndim(forces) else forces)
            out[e, n] = tactile_obs(color, h.bg_img, h.mask)
    return out
