"""Synthetic input sequences of SURVEY.md 8(d), shared by bench.py and the whole-model parity tests (CPU only, torch).

    frames, camera = synth_sequence(n_frames, b, H, W, camera="kitti", seed=1234)

* ``RGB_im [b,H,W,3]``: U[0,1) noise low-pass filtered (3x3 box, two passes) so that features are spatially correlated; frame
  t+1 is frame t seen from the moved camera assuming a fronto-parallel plane at a per-sequence depth U[5,50] m
  (``X_prev = R X_cur + t``, the convention of utils/depth_operations.py:18-53), which keeps the parallax cost volume peaked.
* ``rot [b,4]`` = normalise([1, N(0,0.01)^3]) (w,x,y,z);  ``trans [b,3]`` = N([0,0,1],[0.05,0.05,0.3]) with |t| >= 0.05.
* camera: KITTI-shaped f=(0.580948 W, 1.924101 H), c=(0.490788 W, 0.460944 H) (dataloaders/kitti.py:29-30); Mid-Air-shaped
  f=c=(0.5W,0.5H) (midair.py:20-23); TartanAir-shaped f=(0.5W, 2/3 H), c=(0.5W,0.5H) (tartanair.py:15-18).
"""
import torch

CAMERAS = {
    "kitti": ((0.580948, 1.924101), (0.490788, 0.460944)),
    "midair": ((0.5, 0.5), (0.5, 0.5)),
    "tartan": ((0.5, 2.0 / 3.0), (0.5, 0.5)),
}


def camera_for(kind, b, H, W):
    (fx, fy), (cx, cy) = CAMERAS[kind]
    return {"f": torch.tensor([[fx * W, fy * H]] * b, dtype=torch.float32), "c": torch.tensor([[cx * W, cy * H]] * b, dtype=torch.float32)}


def _rot_mat(q):
    w, x, y, z = q.unbind(-1)
    return torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)), -1).reshape(-1, 3, 3)


def synth_sequence(n_frames, b, H, W, camera="kitti", seed=1234):
    g = torch.Generator().manual_seed(seed)
    cam = camera_for(camera, b, H, W)
    img = torch.rand(b, 3, H, W, generator=g)
    k = torch.ones(3, 1, 3, 3) / 9.0
    for _ in range(2):
        img = torch.nn.functional.conv2d(torch.nn.functional.pad(img, (1, 1, 1, 1), mode="replicate"), k, groups=3)
    depth = 5.0 + 45.0 * torch.rand(b, generator=g)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5, indexing="ij")
    fx, fy = cam["f"][:, 0].view(b, 1, 1), cam["f"][:, 1].view(b, 1, 1)
    cx, cy = cam["c"][:, 0].view(b, 1, 1), cam["c"][:, 1].view(b, 1, 1)
    frames = []
    for t in range(n_frames):
        rot = torch.cat([torch.ones(b, 1), 0.01 * torch.randn(b, 3, generator=g)], 1)
        rot = rot / rot.norm(dim=1, keepdim=True)
        trans = torch.tensor([0.0, 0.0, 1.0]) + torch.randn(b, 3, generator=g) * torch.tensor([0.05, 0.05, 0.3])
        small = trans.norm(dim=1) < 0.05
        trans[small] = torch.tensor([0.0, 0.0, 0.05])
        if t > 0:
            # current pixel -> point on the plane -> previous camera -> previous pixel; sample the previous frame there
            R = _rot_mat(rot)
            d = depth.view(b, 1, 1)
            X = torch.stack(((xs - cx) / fx * d, (ys - cy) / fy * d, d.expand(b, H, W)), -1)          # [b,H,W,3]
            Xp = torch.einsum("bij,bhwj->bhwi", R, X) + trans.view(b, 1, 1, 3)
            u = Xp[..., 0] / Xp[..., 2] * fx + cx
            v = Xp[..., 1] / Xp[..., 2] * fy + cy
            grid = torch.stack((u / W * 2 - 1, v / H * 2 - 1), -1)
            img = torch.nn.functional.grid_sample(img, grid, mode="bilinear", padding_mode="reflection", align_corners=False)
        frames.append({"RGB_im": img.permute(0, 2, 3, 1).contiguous(), "rot": rot.contiguous(), "trans": trans.contiguous()})
    return frames, cam
