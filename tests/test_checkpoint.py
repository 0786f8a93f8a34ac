"""TF tensor-bundle reader / writer (m4depth_b200/checkpoint.py): the format the reference saves its weights in
(callbacks.py:119-129) and ships them in (pretrained_weights.zip).  CPU only; the shipped archive is used when the
reference tree is mounted (build container), never on the GPU box."""
import os

import numpy as np
import pytest
import torch

import oracle

ck = pytest.importorskip("m4depth_b200.checkpoint")
REF_ZIP = "/root/reference/pretrained_weights.zip"


def test_bundle_round_trip_and_checksums(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {
        "a/kernel/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal((3, 3, 5, 7)).astype(np.float32),
        "a/bias/.ATTRIBUTES/VARIABLE_VALUE": rng.standard_normal(7).astype(np.float32),
        "scalar": np.float32(3.5).reshape(()),
        "save_counter/.ATTRIBUTES/VARIABLE_VALUE": np.array(71, dtype=np.int64),
        "z/half": rng.standard_normal((2, 4)).astype(np.float16),
    }
    prefix = str(tmp_path / "cp-0001.ckpt")
    ck.write_bundle(prefix, tensors)
    got = ck.read_bundle(prefix, verify=True)
    assert sorted(got) == sorted(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)
    # a flipped byte in the data file is caught by the per-tensor CRC-32C, one in the index by the block checksum
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[10] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ck.CheckpointError):
        ck.read_bundle(prefix, verify=True)
    ck.read_bundle(prefix, verify=False)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[20] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ck.CheckpointError):
        ck.read_bundle(prefix, verify=True)
    with pytest.raises(ck.CheckpointError):
        ck.read_index(b"\x00" * 100)


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 test vectors
    assert ck.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(b"123456789") == 0xE3069283


def test_model_weights_round_trip_in_reference_key_layout(tmp_path):
    w = oracle.init_weights(3, seed=3, bias_std=0.05, dn_random=True)
    prefix = str(tmp_path / "weights" / "cp-0003.ckpt")
    ck.save_reference_weights(prefix, w)
    names = ck.read_index(open(prefix + ".index", "rb").read())[1]
    assert "encoder/conv_layers_s1/0/kernel/.ATTRIBUTES/VARIABLE_VALUE" in names
    back = ck.load_reference_weights(prefix)
    assert sorted(back) == sorted(w)
    for k in w:
        assert np.array_equal(back[k], np.asarray(w[k]))


@pytest.mark.skipif(not os.path.exists(REF_ZIP), reason="reference tree not mounted")
@pytest.mark.parametrize("which,nparams", [("midair", 4492238), ("kitti", 4492238)])
def test_shipped_checkpoints_load_into_the_oracle_model(which, nparams):
    w = ck.load_reference_weights(REF_ZIP, which)
    ref = oracle.init_weights(6, seed=0)
    assert sorted(w) == sorted(ref)
    assert all(tuple(w[k].shape) == tuple(ref[k].shape) for k in ref)
    assert sum(v.size for v in w.values()) == nparams
    assert all(np.isfinite(v).all() for v in w.values())
    # two frames of the first three pyramid levels with the trained weights on noise images: finite, mostly positive depth
    g = torch.Generator().manual_seed(0)
    model = oracle.M4Depth({k: torch.from_numpy(v) for k, v in w.items()}, nbre_levels=3)
    cam = {"f": torch.tensor([[32.0, 32.0]]), "c": torch.tensor([[32.0, 32.0]])}
    rot = torch.tensor([[1.0, 0.002, -0.001, 0.0015]])
    rot = rot / rot.norm()
    out = None
    for t in range(2):
        s = {"RGB_im": torch.rand(1, 64, 64, 3, generator=g), "rot": rot, "trans": torch.tensor([[0.02, -0.01, 0.8]]), "new_traj": [t == 0]}
        out = model([[s], cam])["depth"]
    assert out.shape == (1, 64, 64, 1) and torch.isfinite(out).all() and float(out.median()) > 0
