"""csv dataloader front-end for the inference path: the reference's ``dataloaders/`` (generic.py:84-145, kitti.py, midair.py,
tartanair.py) without TensorFlow, for ``usecase`` "eval" / "predict" (what feeds ``M4Depth.test_step`` / ``predict_step``).

Same names, arguments and sample layout as the reference:

    loader = get_loader("kitti-raw")                                   # dataloaders/__init__.py:6-17
    settings = DataloaderParameters(db_path_config, records_path, db_seq_len, seq_len, augment)   # generic.py:8
    dataset = loader.get_dataset("eval", settings, batch_size=1, out_size=[256, 768])
    for sample in dataset: model.test_step(sample)                     # loader.length samples

A sample is a dict of torch CPU tensors (pinned when CUDA is available, so ``M4Depth.call`` uploads them asynchronously):
``RGB_im [b,H,W,3]`` in [0,1], ``depth [b,H,W,1]`` (when the csv has a depth column), ``rot [b,4]`` (w,x,y,z), ``trans [b,3]``,
``new_traj [b]`` (csv ``id == 0``), ``camera {"f": [b,2], "c": [b,2]}``; with ``db_seq_len`` the per-frame entries gain a time
axis (``[b,T,...]``, ``new_traj [b,T]`` true at t = 0, camera of the first frame: generic.py:160-186) - the KITTI protocol that
``M4Depth.test_step`` scores on the last frame.  Records are tab-separated csv files (one trajectory each) found recursively
under ``records_path``; image paths are relative to ``db_path_config[dataset_name]``.

Not carried over: "train" / "finetune" (shuffling, random cuts, colour / flip augmentation - the training loop is out of scope)
and the tf.data prefetch machinery; images are decoded with Pillow.  Resizing follows TF2 ``tf.image.resize``: bilinear with
half-pixel centres and no antialiasing for colour / dense depth, half-pixel NEAREST for sparse depth.
"""
import glob
import os
from collections import namedtuple

import numpy as np
import torch

DataloaderParameters = namedtuple('DataloaderParameters', ('db_path_config', 'records_path', 'db_seq_len', 'seq_len', 'augment'))


def _resize_bilinear(img, out_size):
    """tf.image.resize(img, out_size) (bilinear, half-pixel centres, antialias=False); img [h,w,c] float32."""
    t = torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1).unsqueeze(0)
    t = torch.nn.functional.interpolate(t, size=tuple(out_size), mode="bilinear", align_corners=False, antialias=False)
    return t[0].permute(1, 2, 0).contiguous()


def _resize_nearest(img, out_size):
    """tf.image.resize(img, out_size, method='nearest') (TF2: src = min(floor((dst + 0.5) * in / out), in - 1))."""
    h, w = img.shape[:2]
    ys = np.minimum(np.floor((np.arange(out_size[0], dtype=np.float32) + 0.5) * np.float32(h / out_size[0])).astype(np.int64), h - 1)
    xs = np.minimum(np.floor((np.arange(out_size[1], dtype=np.float32) + 0.5) * np.float32(w / out_size[1])).astype(np.int64), w - 1)
    return torch.from_numpy(np.ascontiguousarray(img[ys][:, xs]))


def _read_image(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im)


class Dataset:
    """What get_dataset returns: an iterable of batched samples with ``cardinality()`` / ``len()``."""

    def __init__(self, loader, items, batch_size, sequences):
        self._loader, self._items, self._bs, self._seq = loader, items, batch_size, sequences

    def __len__(self):
        return len(self._items) // self._bs

    def cardinality(self):
        return len(self)

    def __iter__(self):
        pin = torch.cuda.is_available()
        for i in range(len(self)):
            group = self._items[i * self._bs:(i + 1) * self._bs]
            samples = [self._loader._build_sequence_samples(g) if self._seq else self._loader._decode_samples(g) for g in group]
            yield _collate(samples, pin)


def _collate(samples, pin):
    out = {}
    for k in samples[0]:
        if isinstance(samples[0][k], dict):
            out[k] = _collate([s[k] for s in samples], pin)
        else:
            t = torch.stack([torch.as_tensor(s[k]) for s in samples], 0)
            out[k] = t.pin_memory() if pin and t.is_floating_point() else t
    return out


class DataLoaderGeneric:
    """Superclass of the dataset loaders (generic.py:10-145), inference use cases."""

    def __init__(self, dataset_name):
        self.build_functions = {"eval": self._build_eval_dataset, "predict": self._build_eval_dataset}
        self.augment = None
        self.settings = None
        self.db_name = dataset_name

    def _decode_samples(self, data_sample):
        raise NotImplementedError

    def _set_output_size(self, out_size=None):
        raise NotImplementedError

    def get_dataset(self, usecase, settings, batch_size=3, out_size=None):
        if out_size is None:
            self._set_output_size()
        else:
            self._set_output_size(out_size=list(out_size))
        self.settings = settings
        self.records_path = settings.records_path
        self.db_path = settings.db_path_config[self.db_name]
        self.db_seq_len = settings.db_seq_len
        self.seq_len = settings.seq_len
        self.batch_size = batch_size
        self.usecase = usecase
        if not (self.db_seq_len is None or self.seq_len is None) and self.db_seq_len < self.seq_len:
            raise Exception('db_seq_len must be larger or equal than seq_len')
        try:
            function = self.build_functions[usecase]
        except KeyError:
            raise Exception('Usecase "%s" not implemented for this dataloader (inference front-end: eval / predict)' % usecase)
        self.dataset = function()
        self.length = self.dataset.cardinality()
        return self.dataset

    def _get_trajectories(self):
        import pandas as pd
        csv_files = sorted(glob.glob(os.path.join(self.records_path, "**/*.csv"), recursive=True))
        trajectories = [pd.read_csv(f, sep="\t").to_dict("records") for f in csv_files]
        if trajectories == []:
            raise Exception("No csv files found at the given path: %s" % self.records_path)
        return trajectories

    def _build_eval_dataset(self):
        """generic.py:124-145: frames in trajectory order, one per batch; or, with db_seq_len, consecutive sub-sequences of that
        length (remainder dropped) batched by batch_size."""
        self.augment = False
        trajectories = self._get_trajectories()
        if self.db_seq_len is None:
            items = [rec for traj in trajectories for rec in traj]
            return Dataset(self, items, 1, sequences=False)
        self.seq_len = self.db_seq_len
        items = []
        for traj in trajectories:
            for s in range(len(traj) // self.db_seq_len):
                items.append(traj[s * self.db_seq_len:(s + 1) * self.db_seq_len])
        return Dataset(self, items, self.batch_size, sequences=True)

    def _build_sequence_samples(self, records):
        """generic.py:160-186: stack the frames of a sub-sequence; new_traj is true at its first frame; the camera is the
        first frame's."""
        frames = [self._decode_samples(r) for r in records]
        out = {"camera": dict(frames[0]["camera"])}
        for k in ("depth", "RGB_im", "rot", "trans"):
            if k in frames[0]:
                out[k] = torch.stack([f[k] for f in frames], 0)
        out["new_traj"] = torch.tensor([i == 0 for i in range(len(frames))])
        return out

    # shared by the three datasets
    def _pose(self, r):
        return (torch.tensor([r['qw'], r['qx'], r['qy'], r['qz']], dtype=torch.float32),
                torch.tensor([r['tx'], r['ty'], r['tz']], dtype=torch.float32), torch.tensor(int(r['id']) == 0))

    def _rgb(self, r, size):
        img = _read_image(os.path.join(self.db_path, r['camera_l'])).astype(np.float32) / np.float32(255.)
        if img.ndim == 2:
            img = np.repeat(img[..., None], 3, -1)
        return _resize_bilinear(img[..., :3], size)


class DataLoaderKittiRaw(DataLoaderGeneric):
    """dataloaders/kitti.py: intrinsics from the csv (fractions of the image size), velodyne depth in uint16 PNG / 256,
    nearest resize, Garg / Eigen evaluation crop."""

    def __init__(self):
        super().__init__('kitti-raw')
        self.in_size = [370, 1220]
        self.depth_type = "velodyne"

    def _set_output_size(self, out_size=[256, 768]):
        self.out_size = list(out_size)
        crop = np.array([0.40810811 * out_size[0], 0.99189189 * out_size[0],
                         0.03594771 * out_size[1], 0.96405229 * out_size[1]]).astype(np.int32)
        crop_mask = np.zeros(self.out_size + [1], dtype=np.float32)
        crop_mask[crop[0]:crop[1], crop[2]:crop[3], :] = 1
        self.eval_crop_mask = torch.from_numpy(crop_mask)

    def _decode_samples(self, r):
        H, W = self.out_size
        rot, trans, new_traj = self._pose(r)
        out = {"camera": {"f": torch.tensor([r['fx'] * W, r['fy'] * H], dtype=torch.float32),
                          "c": torch.tensor([r['cx'] * W, r['cy'] * H], dtype=torch.float32)},
               "RGB_im": self._rgb(r, self.out_size), "rot": rot, "trans": trans, "new_traj": new_traj}
        if 'depth' in r:
            d = _read_image(os.path.join(self.db_path, r['depth'])).astype(np.float32) / np.float32(256)
            depth = _resize_nearest(d.reshape(d.shape[0], d.shape[1], 1), self.out_size)
            out['depth'] = depth * self.eval_crop_mask if self.usecase == "eval" else depth
        return out


class DataLoaderMidAir(DataLoaderGeneric):
    """dataloaders/midair.py: fixed intrinsics f = c = size / 2, depth = 512 / disparity stored as float16 bits in a uint16 PNG."""

    def __init__(self, out_size=[384, 384], crop=False):
        super().__init__('midair')
        self.in_size = [1024, 1024]
        self.depth_type = "map"
        self.crop = False

    def _set_output_size(self, out_size=[384, 384]):
        self.out_size = list(out_size)
        self.intermediate_size = self.out_size
        self.fx, self.fy = 0.5 * self.out_size[1], 0.5 * self.out_size[0]
        self.cx, self.cy = 0.5 * self.out_size[1], 0.5 * self.out_size[0]

    def get_dataset(self, usecase, settings, batch_size=3, out_size=[384, 384], crop=False):
        if crop:
            raise AttributeError("Crop option should be disabled when evaluating or predicting samples")
        return super().get_dataset(usecase, settings, batch_size=batch_size, out_size=out_size)

    def _decode_samples(self, r):
        rot, trans, new_traj = self._pose(r)
        out = {"camera": {"f": torch.tensor([self.fx, self.fy], dtype=torch.float32), "c": torch.tensor([self.cx, self.cy], dtype=torch.float32)},
               "RGB_im": self._rgb(r, self.intermediate_size), "rot": rot, "trans": trans, "new_traj": new_traj}
        if 'disp' in r:
            raw = _read_image(os.path.join(self.db_path, r['disp'])).astype(np.uint16)
            depth = np.float32(512.) / raw.view(np.float16).astype(np.float32)
            out['depth'] = _resize_bilinear(depth.reshape(depth.shape[0], depth.shape[1], 1), self.intermediate_size)
        return out


class DataLoaderTartanAir(DataLoaderGeneric):
    """dataloaders/tartanair.py: fixed intrinsics, depth as the trailing h*w float32 values of a .npy file, nearest resize,
    masked where the colour image is black."""

    def __init__(self, out_size=[384, 512]):
        super().__init__('tartanair')
        self.in_size = [480, 640]
        self.depth_type = "map"

    def _set_output_size(self, out_size=[384, 512]):
        self.out_size = list(out_size)
        self.fx, self.fy = 0.5 * self.out_size[1], 2. / 3. * self.out_size[0]
        self.cx, self.cy = 0.5 * self.out_size[1], 0.5 * self.out_size[0]

    def _decode_samples(self, r):
        rot, trans, new_traj = self._pose(r)
        rgb = self._rgb(r, self.out_size)
        out = {"camera": {"f": torch.tensor([self.fx, self.fy], dtype=torch.float32), "c": torch.tensor([self.cx, self.cy], dtype=torch.float32)},
               "RGB_im": rgb, "rot": rot, "trans": trans, "new_traj": new_traj}
        if 'depth' in r:
            mask = (rgb.pow(2).sum(-1, keepdim=True).sqrt() > 0).to(torch.float32)
            with open(os.path.join(self.db_path, r['depth']), "rb") as f:      # .npy: header, then h*w float32 (tartanair.py:40-42)
                raw = f.read()
            depth = np.frombuffer(raw[-(self.in_size[0] * self.in_size[1] * 4):], dtype=np.float32).reshape(self.in_size + [1])
            out['depth'] = _resize_nearest(depth, self.out_size) * mask
        return out


MidAir, KittiRaw, TartanAir = DataLoaderMidAir, DataLoaderKittiRaw, DataLoaderTartanAir


def get_loader(name):
    """dataloaders/__init__.py:6-17."""
    available = {"midair": MidAir, "kitti-raw": KittiRaw, "tartanair": TartanAir}
    try:
        return available[name]()
    except KeyError:
        print("Dataloaders available:")
        print(available.keys())
        raise NotImplementedError
