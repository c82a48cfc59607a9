"""CPU oracle for the (P) point-cloud path — TEST INFRASTRUCTURE ONLY.

Restates, with CPU torch ops in the reference's own order and serial per-env
structure (so the `torch.randint` stream matches):
  PointCloudGenerator.__init__/convert   isaacgyminsertion/tasks/utils/pcl_utils.py:28-90
  CameraPointCloud.get_point_cloud       pcl_utils.py:168-184
  CameraPointCloud.sample_n              pcl_utils.py:195-201
  filter_pts                             tasks/factory_tactile/factory_task_insertion.py:66-77
  seg masking + assembly                 factory_task_insertion.py:956-989,1014-1027

Parity pin: tests/golden/pcl_golden.npz holds outputs of the REAL reference classes
(imported from /root/reference with isaacgym/matplotlib stubbed, tools/make_golden_pcl.py);
tests/test_oracle_pcl.py checks this restatement against them bit-for-bit.
"""
import numpy as np
import torch


def filter_pts(pts):
    """factory_task_insertion.py:66-77."""
    x = pts[:, 0]
    y = pts[:, 1]
    z = pts[:, 2]
    valid1 = (z >= 0.001) & (z <= 0.6)
    valid2 = (x >= 0.1) & (x <= 0.7)
    valid3 = (y >= -0.4) & (y <= 0.4)
    return pts[valid1 & valid3 & valid2]


class CameraOracle:
    """One env's camera: pcl_utils.py:29-60 (construction) and :62-90 (convert)."""

    def __init__(self, proj_matrix, view_matrix, env_to_global, width, height, depth_max=1.0):
        fu = 2 / proj_matrix[0, 0]
        fv = 2 / proj_matrix[1, 1]
        self.fu = width / fu
        self.fv = height / fv
        self.cu = width / 2.
        self.cv = height / 2.
        self.int_mat = torch.Tensor([[-self.fu, 0, self.cu], [0, self.fv, self.cv], [0, 0, 1]])
        self.ext_mat = torch.inverse(torch.Tensor(view_matrix))
        self.int_mat_T_inv = torch.inverse(self.int_mat.T)
        self.depth_max = depth_max
        self.env_to_global = env_to_global
        x, y = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
        uv_one = torch.stack((y, x, torch.ones_like(x)), dim=-1).float()
        self.uv_one_in_cam = uv_one @ self.int_mat_T_inv

    def convert(self, depth_buffer):
        if self.depth_max is not None:
            valid_ids = depth_buffer > -self.depth_max
        else:
            valid_ids = torch.ones(depth_buffer.shape, dtype=bool)
        valid_depth = depth_buffer[valid_ids]
        uv = self.uv_one_in_cam[valid_ids]
        pts_in_cam = torch.mul(uv, valid_depth.unsqueeze(-1))
        pts_in_cam = torch.cat((pts_in_cam, torch.ones(*pts_in_cam.shape[:-1], 1)), dim=-1)
        pts_in_world = pts_in_cam @ self.ext_mat
        e2g = torch.inverse(torch.Tensor(self.env_to_global))
        pts_in_world = torch.matmul(pts_in_world, e2g.T)
        return pts_in_world[..., :3]


def build_cameras(gym, depth_max=1.0):
    cams = []
    for e in gym.envs:
        o = gym.get_env_origin(e)
        e2g = np.identity(4)
        e2g[:3, 3] = np.array([o.x, o.y, o.z])
        cams.append(CameraOracle(gym.get_camera_proj_matrix(None, e, e), gym.get_camera_view_matrix(None, e, e),
                                 e2g, gym.width, gym.height, depth_max))
    return cams


def get_ptd(cams, depths, filter_func=filter_pts):
    """pcl_utils.py:203-212 (+ _proc_pts :186-193)."""
    out = []
    for e, cam in enumerate(cams):
        pts = cam.convert(depths[e])
        if filter_func is not None:
            pts = filter_func(pts)
        out.append(pts)
    return out


def get_point_cloud(cams, depths, sample_num, filter_func=filter_pts, return_idx=False):
    """pcl_utils.py:168-184: zeros unless `pts.any()`, else pts[torch.randint(0, M, (m,))]."""
    n = len(cams)
    out = torch.zeros((n, sample_num, 3))
    idx = torch.zeros((n, sample_num), dtype=torch.int64)
    all_pts = get_ptd(cams, depths, filter_func)
    for e in range(n):
        if all_pts[e].any():
            ids = torch.randint(0, all_pts[e].shape[0], size=(sample_num,))
            out[e] = all_pts[e][ids]
            idx[e] = ids
    return (out, idx, all_pts) if return_idx else out


def masked_depth(depth, seg, seg_id):
    """factory_task_insertion.py:956-959 / :975."""
    return (depth.flatten(start_dim=1) * (seg.flatten(start_dim=1) == seg_id)).reshape(depth.shape)


def pcl_observation(cams, depth, seg, num_points=400, num_points_socket=400, include_all_pcl=False, total_points=2048):
    """[all-scene cloud first when include_all_pcl (:946-949)], plug cloud, then socket cloud (RNG order of
    factory_task_insertion.py:946-979), merged in pcl_components order plug | socket | all (:1014-1027)."""
    all_pts = get_point_cloud(cams, depth, total_points) if include_all_pcl else None
    plug = get_point_cloud(cams, masked_depth(depth, seg, 2), num_points)
    socket = get_point_cloud(cams, masked_depth(depth, seg, 3), num_points_socket)
    parts = [plug, socket] + ([all_pts] if include_all_pcl else [])
    return torch.cat(parts, dim=1).flatten(start_dim=1)
