"""CPU: oracle/student.py against the golden outputs of the REAL reference code
(tests/golden/student_golden.npz, tools/make_golden_student.py) and the Random123 known answers."""
import os

import numpy as np
import pytest
import torch

from oracle import student as ost


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "student_golden.npz"))


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    def run(ctr, key):
        idx = np.array([ctr[0] | (ctr[1] << 32)], dtype=np.uint64)
        return [int(v) for v in ost.philox(idx, ctr[2], ctr[3], key[0] | (key[1] << 32))[0]]
    assert run((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draw_statistics():
    u = ost.philox_uniform_image(64, 5184, 1, 0, 123)
    assert 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 2e-3 and abs(u.var() - 1 / 12) < 1e-3
    z, ug = ost.philox_pcl_draws(256, 400, 1, 123)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs(np.mean(z ** 4) - 3.0) < 0.1          # kurtosis of a normal
    assert abs(np.corrcoef(z[..., 0].ravel(), z[..., 1].ravel())[0, 1]) < 0.01
    assert abs((ug < 0.3).mean() - 0.3) < 5e-3


def test_cam_image_obs_matches_reference(g):
    N, H, W = g["depth"].shape
    npix = H * W
    seed, step = int(g["seed"]), int(g["step"])
    u_d = torch.from_numpy(ost.philox_uniform_image(N, npix, step, 0, seed))
    u_s = torch.from_numpy(ost.philox_uniform_image(N, npix, step, 1, seed))
    image_buf = torch.full((N, npix), -7.0)
    seg_buf = torch.full((N, npix), -7, dtype=torch.int32)
    ost.cam_image_obs(torch.from_numpy(g["depth"]), torch.from_numpy(g["seg"]), torch.from_numpy(g["update"]),
                      torch.from_numpy(g["update_seg"]), torch.from_numpy(g["seg_noise"]), image_buf, seg_buf, u_d, u_s,
                      float(g["dis_noise"]), float(g["far_clip"]), float(g["near_clip"]), float(g["flip_prob"]))
    np.testing.assert_array_equal(image_buf.numpy(), g["image_buf"])   # NaNs compare equal here
    np.testing.assert_array_equal(seg_buf.numpy(), g["seg_buf"])
    # the fixture exercises what it claims to
    assert (g["seg_buf"][g["update_seg"] & g["seg_noise"]] != g["seg"].reshape(N, -1)[g["update_seg"] & g["seg_noise"]]).any()
    assert (g["image_buf"][~g["update"]] == -7.0).all() and np.isnan(g["image_buf"]).any() == bool(g["update"][3])


def test_random_noise_matches_reference(g):
    B, P, _ = g["pts"].shape
    z, u = ost.philox_pcl_draws(B, P, int(g["step"]), int(g["seed"]))
    out = ost.random_noise(torch.from_numpy(g["pts"]), torch.from_numpy(g["pcl_pos_noise"]), torch.from_numpy(z),
                           torch.from_numpy(u))
    np.testing.assert_array_equal(out.numpy(), g["noisy"])
    assert np.abs(g["noisy"] - g["pts"]).max() <= 2e-3 + 1e-7


def test_running_mean_std_matches_reference(g):
    m = ost.RunningMeanStd(3)
    for b, want in zip(g["rms_batches"][:3], g["rms_out"]):
        np.testing.assert_array_equal(m(torch.from_numpy(b)).numpy(), want)
    np.testing.assert_array_equal(m.running_mean.numpy(), g["rms_mean"])
    np.testing.assert_array_equal(m.running_var.numpy(), g["rms_var"])
    np.testing.assert_array_equal(m.count.numpy(), g["rms_count"])
    m.training = False
    ev = m(torch.from_numpy(g["rms_batches"][3]))
    np.testing.assert_array_equal(ev.numpy(), g["rms_eval"])
    np.testing.assert_array_equal(m(ev, unnorm=True).numpy(), g["rms_unnorm"])


def test_process_obs_matches_reference(g):
    obs = {"student_obs": torch.from_numpy(g["po_stud_in"]), "img": torch.from_numpy(g["po_img_in"]),
           "seg": torch.from_numpy(g["po_seg_in"]), "pcl": torch.from_numpy(g["po_pcl_in"])}
    out = ost.process_obs(obs, ost.RunningMeanStd(3), ost.RunningMeanStd(18))
    for k, gk in (("seg", "po_seg"), ("img", "po_img"), ("pcl", "po_pcl"), ("student_obs", "po_stud")):
        np.testing.assert_array_equal(out[k].numpy(), g[gk])
    nd = ost.process_obs({"img": obs["img"], "seg": obs["seg"]}, distinct=False)
    np.testing.assert_array_equal(nd["seg"].numpy(), g["po_nd_seg"])
    np.testing.assert_array_equal(nd["img"].numpy(), g["po_nd_img"])


def test_queue_push():
    q = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)
    q0 = q.clone()
    x = torch.full((2, 4), -1.0)
    ost.queue_push(q, x)
    assert torch.equal(q[:, 0], x) and torch.equal(q[:, 1:], q0[:, :-1])


def test_tactile_transform_equals_per_image_loop():
    """The batched TactileTransform == the reference's per-image loop (utils.py:140-156) for its eval transform
    (Resize bilinear -> CenterCrop, utils.py:217-274), and is skipped at the identity size."""
    from torchvision import transforms
    from isaacgyminsertion_b200.student_obs import TactileTransform
    g = torch.Generator().manual_seed(0)
    x = torch.rand((2, 2, 3, 1, 32, 64), generator=g)

    def loop(tf, tac):                       # the reference's __call__
        B, T, F, C, H, W = tac.shape
        flat = tac.view(-1, C, H, W)
        out = torch.stack([tf(flat[i]) for i in range(flat.shape[0])])
        return out.view(B, T, F, C, *out.shape[2:])
    ident = torch.nn.Sequential(transforms.Resize((32, 64), interpolation=transforms.InterpolationMode.BILINEAR),
                                transforms.CenterCrop((32, 64)))
    assert torch.equal(loop(ident, x), x)                                     # identity at the shipped sizes
    assert TactileTransform(ident, identity_size=(32, 64))(x) is x
    assert torch.equal(TactileTransform(ident)(x), x)
    small = torch.nn.Sequential(transforms.Resize((16, 32), interpolation=transforms.InterpolationMode.BILINEAR),
                                transforms.CenterCrop((12, 24)))
    got, want = TactileTransform(small)(x), loop(small, x)
    assert got.shape == (2, 2, 3, 1, 12, 24) and torch.allclose(got, want, atol=1e-6)
