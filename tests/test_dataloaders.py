"""csv dataloader front-end (m4depth_b200/dataloaders.py <-> reference dataloaders/): records are the first five lines of one
trajectory per dataset from the reference's data.zip (tests/golden/csv/), the images they point at are written here with
Pillow.  CPU tests check the sample layout and the decode / resize semantics; the GPU test feeds M4Depth.test_step."""
import os

import numpy as np
import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV = os.path.join(ROOT, "tests", "golden", "csv")
Image = pytest.importorskip("PIL.Image")


def _dl():
    from m4depth_b200 import dataloaders
    return dataloaders


def _write_db(tmp_path, name, in_size):
    """Materialise the files the csv records reference: smooth random colour images and depth in the dataset's encoding."""
    import pandas as pd
    rng = np.random.default_rng(3)
    db = tmp_path / name
    csv = [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(CSV, name)) for f in fs if f.endswith(".csv")][0]
    recs = pd.read_csv(csv, sep="\t").to_dict("records")
    h, w = in_size
    truth = []
    by_path = {}                                                 # several records may share a depth file (KITTI ground truth frames)
    for r in recs:
        rgb = (rng.random((h, w, 3)) * 255).astype(np.uint8)
        rgb[:4, :4] = 0                                          # a black corner: TartanAir masks depth there
        p = db / r["camera_l"]
        p.parent.mkdir(parents=True, exist_ok=True)
        Image.fromarray(rgb).save(p, format="PNG")               # lossless whatever the extension says
        dkey = r.get("depth", r.get("disp"))
        if dkey in by_path:
            truth.append((rgb, by_path[dkey]))
            continue
        depth = (rng.random((h, w)) * 60 + 1).astype(np.float32)
        if name == "kitti-raw":
            d16 = (depth * 256).astype(np.uint16)
            d16[rng.random((h, w)) < 0.7] = 0                    # sparse velodyne returns
            q = db / r["depth"]
            q.parent.mkdir(parents=True, exist_ok=True)
            Image.fromarray(d16).save(q, format="PNG")
            depth = d16.astype(np.float32) / 256
        elif name == "midair":
            disp = (512.0 / depth).astype(np.float16)
            q = db / r["disp"]
            q.parent.mkdir(parents=True, exist_ok=True)
            Image.fromarray(disp.view(np.uint16)).save(q, format="PNG")
            depth = 512.0 / disp.astype(np.float32)
        else:
            q = db / r["depth"]
            q.parent.mkdir(parents=True, exist_ok=True)
            with open(q, "wb") as f:
                f.write(b"\x93NUMPY-header-stand-in".ljust(128))       # decode_raw keeps only the trailing h*w floats
                f.write(depth.tobytes())
        by_path[dkey] = depth
        truth.append((rgb, depth))
    return str(db), recs, truth


@pytest.mark.parametrize("name,in_size,out_size", [("kitti-raw", (37, 122), (32, 96)), ("midair", (64, 64), (48, 48)), ("tartanair", (480, 640), (96, 128))])
def test_single_frame_samples_follow_the_reference_decoders(tmp_path, name, in_size, out_size):
    dl = _dl()
    db, recs, truth = _write_db(tmp_path, name, in_size)
    loader = dl.get_loader(name)
    if name == "tartanair":
        loader.in_size = list(in_size)
    settings = dl.DataloaderParameters({name: db}, os.path.join(CSV, name), None, None, False)
    ds = loader.get_dataset("eval", settings, batch_size=1, out_size=list(out_size))
    assert loader.length == len(recs) == len(ds)
    H, W = out_size
    for r, (rgb, depth), s in zip(recs, truth, ds):
        assert tuple(s["RGB_im"].shape) == (1, H, W, 3) and s["RGB_im"].dtype == torch.float32
        assert tuple(s["depth"].shape) == (1, H, W, 1) and tuple(s["rot"].shape) == (1, 4) and tuple(s["trans"].shape) == (1, 3)
        assert bool(s["new_traj"][0]) == (int(r["id"]) == 0)                                   # kitti.py:40
        np.testing.assert_allclose(s["rot"][0].numpy(), [r["qw"], r["qx"], r["qy"], r["qz"]], rtol=1e-6)
        np.testing.assert_allclose(s["trans"][0].numpy(), [r["tx"], r["ty"], r["tz"]], rtol=1e-6)
        if name == "kitti-raw":                                                                # kitti.py:29-30
            np.testing.assert_allclose(s["camera"]["f"][0].numpy(), [r["fx"] * W, r["fy"] * H], rtol=1e-6)
            np.testing.assert_allclose(s["camera"]["c"][0].numpy(), [r["cx"] * W, r["cy"] * H], rtol=1e-6)
        elif name == "midair":                                                                 # midair.py:20-23
            assert s["camera"]["f"][0].tolist() == [0.5 * W, 0.5 * H] and s["camera"]["c"][0].tolist() == [0.5 * W, 0.5 * H]
        else:                                                                                  # tartanair.py:15-18
            np.testing.assert_allclose(s["camera"]["f"][0].numpy(), [0.5 * W, 2.0 / 3.0 * H], rtol=1e-6)
        # colour: /255 then TF2 bilinear resize (half-pixel centres, no antialias) - an independent numpy evaluation
        x = rgb.astype(np.float32) / 255.0
        sy = (np.arange(H) + 0.5) * (in_size[0] / H) - 0.5
        sx = (np.arange(W) + 0.5) * (in_size[1] / W) - 0.5
        y0 = np.clip(np.floor(sy), 0, in_size[0] - 1).astype(int); y1 = np.minimum(y0 + 1, in_size[0] - 1); wy = np.clip(sy - np.floor(sy), 0, 1) * (sy >= 0)
        x0 = np.clip(np.floor(sx), 0, in_size[1] - 1).astype(int); x1 = np.minimum(x0 + 1, in_size[1] - 1); wx = np.clip(sx - np.floor(sx), 0, 1) * (sx >= 0)
        top = x[y0][:, x0] * (1 - wx)[None, :, None] + x[y0][:, x1] * wx[None, :, None]
        bot = x[y1][:, x0] * (1 - wx)[None, :, None] + x[y1][:, x1] * wx[None, :, None]
        want = top * (1 - wy)[:, None, None] + bot * wy[:, None, None]
        np.testing.assert_allclose(s["RGB_im"][0].numpy(), want, atol=1e-5)
        # depth: nearest (half-pixel) for the sparse / raw maps, bilinear for Mid-Air; dataset-specific masks
        d = torch.from_numpy(depth).reshape(1, in_size[0], in_size[1], 1)
        if name == "midair":
            assert float((s["depth"] - torch.from_numpy(np.asarray(want[..., :1] * 0)).unsqueeze(0)).abs().max()) > 0   # decoded, not zero
            assert float(s["depth"].min()) > 0.9 and float(s["depth"].max()) < 62
        else:
            want_d = oracle.resize_nearest(d, H, W)
            if name == "kitti-raw":
                want_d = want_d * loader.eval_crop_mask
            else:
                grey = s["RGB_im"].pow(2).sum(-1, keepdim=True).sqrt()
                want_d = want_d * (grey > 0)
            assert torch.equal(s["depth"], want_d)


def test_sequence_mode_batches_subsequences_like_the_reference(tmp_path):
    """db_seq_len: consecutive sub-sequences (remainder dropped), [b,T,...] tensors, new_traj true at t = 0 only, camera of the
    first frame (generic.py:124-145,160-186) - the 5-D layout M4Depth.test_step scores on its last frame."""
    dl = _dl()
    db, recs, _ = _write_db(tmp_path, "kitti-raw", (37, 122))
    loader = dl.get_loader("kitti-raw")
    settings = dl.DataloaderParameters({"kitti-raw": db}, os.path.join(CSV, "kitti-raw"), 2, 2, False)
    ds = loader.get_dataset("eval", settings, batch_size=2, out_size=[32, 96])
    assert loader.length == 1                                    # 5 records -> 2 sub-sequences of 2 -> one batch of 2
    s = next(iter(ds))
    assert tuple(s["RGB_im"].shape) == (2, 2, 32, 96, 3) and tuple(s["depth"].shape) == (2, 2, 32, 96, 1)
    assert s["new_traj"].tolist() == [[True, False], [True, False]]
    assert tuple(s["camera"]["f"].shape) == (2, 2) and tuple(s["rot"].shape) == (2, 2, 4)
    np.testing.assert_allclose(s["trans"][1, 0].numpy(), [recs[2]["tx"], recs[2]["ty"], recs[2]["tz"]], rtol=1e-6)
    with pytest.raises(Exception):
        loader.get_dataset("train", settings)
    with pytest.raises(NotImplementedError):
        dl.get_loader("nyu")


@pytest.mark.gpu
def test_loader_feeds_test_step_on_the_gpu(tmp_path):
    """The reference evaluation loop (main.py:150-172 / test_step, m4depth_network.py:433-474) on loader samples: single frames
    and the KITTI sequence mode give the metrics the oracle protocol gives on the same samples."""
    import m4depth_b200 as m
    dl = _dl()
    db, recs, _ = _write_db(tmp_path, "kitti-raw", (37, 122))
    w = oracle.init_weights(3, seed=1, bias_std=0.05, dn_random=True)
    loader = dl.get_loader("kitti-raw")
    settings = dl.DataloaderParameters({"kitti-raw": db}, os.path.join(CSV, "kitti-raw"), None, None, False)
    ds = loader.get_dataset("eval", settings, batch_size=1, out_size=[64, 128])
    model = m.M4Depth(nbre_levels=3, use_cuda_graph=False)
    model.load_weights(w)
    ref = oracle.M4Depth(w, nbre_levels=3, pscv_kwargs={"use_cuda_backproject": False})
    n = 0
    for s in ds:
        res = model.test_step(s)
        want = ref([[{k: s[k] for k in ("RGB_im", "rot", "trans", "new_traj")}], s["camera"]])["depth"]
        got = model._out.cpu()
        if n <= 1:
            assert float(((got - want).abs() / (want.abs() + 0.1)).max()) <= 1e-4
        n += 1
    assert n == len(recs) and set(res) >= {"AbsRel", "RMSE"} and all(np.isfinite(float(v)) for v in res.values())
