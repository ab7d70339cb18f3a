"""Data path: PLY reader / writer and the dataset of graphVAE_train.py:76-123, on the CPU."""
import numpy as np
import pytest
import torch

from vcmeshconv_b200 import ply_dataset as P
from vcmeshconv_b200 import synthetic


def _mesh(n=200, seed=0):
    pts, faces = synthetic.sphere(n, seed=seed)
    return (pts * 1000.0).astype(np.float32), faces          # the reference's PLYs are in millimetres


@pytest.mark.parametrize("binary", [True, False])
def test_ply_roundtrip(tmp_path, binary):
    v, f = _mesh()
    fn = str(tmp_path / "m.ply")
    P.write_ply(fn, v, f, binary=binary)
    got = P.read_ply(fn)
    assert got["vertices"].dtype == np.float32 and got["faces"].dtype == np.int64
    assert np.array_equal(got["faces"], f)
    if binary:
        assert np.array_equal(got["vertices"], v)
    else:
        assert np.allclose(got["vertices"], v, rtol=1e-7, atol=0)
    assert torch.equal(P.load_template_faces(fn), torch.from_numpy(f.astype(np.int64)))


def test_ply_with_extra_properties_and_double_coordinates(tmp_path):
    """Scanner exports carry normals / colours and sometimes doubles: they are skipped / converted."""
    v, f = _mesh(50)
    rec = np.empty(len(v), dtype=[("x", "<f8"), ("nx", "<f4"), ("y", "<f8"), ("z", "<f8"), ("red", "u1")])
    rec["x"], rec["y"], rec["z"], rec["nx"], rec["red"] = v[:, 0], v[:, 1], v[:, 2], 0.5, 7
    faces = np.empty(len(f), dtype=[("n", "u1"), ("v", "<u4", (3,))])
    faces["n"], faces["v"] = 3, f
    hdr = ("ply\nformat binary_little_endian 1.0\ncomment made by a test\nelement vertex %d\nproperty double x\n"
           "property float nx\nproperty double y\nproperty double z\nproperty uchar red\nelement face %d\n"
           "property list uchar uint vertex_index\nend_header\n" % (len(v), len(f)))
    fn = tmp_path / "x.ply"
    fn.write_bytes(hdr.encode() + rec.tobytes() + faces.tobytes())
    got = P.read_ply(str(fn))
    assert np.array_equal(got["vertices"], v) and np.array_equal(got["faces"], f)


def test_bad_files_raise(tmp_path):
    (tmp_path / "a.ply").write_bytes(b"solid not a ply\n")
    with pytest.raises(ValueError):
        P.read_ply(str(tmp_path / "a.ply"))
    v, f = _mesh(30)
    P.write_ply(str(tmp_path / "b.ply"), v, f)
    raw = (tmp_path / "b.ply").read_bytes()
    (tmp_path / "c.ply").write_bytes(raw[:-7])
    with pytest.raises(ValueError):
        P.read_ply(str(tmp_path / "c.ply"))
    (tmp_path / "d.ply").write_bytes(raw.replace(b"binary_little_endian", b"binary_big_endian"))
    with pytest.raises(ValueError):
        P.read_ply(str(tmp_path / "d.ply"))


def test_dataset_matches_reference_formulas(tmp_path):
    v, f = _mesh(120)
    track = str(tmp_path) + "/"
    rng = np.random.default_rng(0)
    frames = [499, 500, 731, 4000, 4001]
    for fr in frames[:-1] + [4001]:
        P.write_ply("%s%06d.ply" % (track, fr), v + rng.standard_normal(v.shape).astype(np.float32), f)
    ds = P.MeshDataset(track, [str(x) for x in frames], f)
    assert [fr for fr, _ in ds.framelist] == ["500", "731", "4000"]           # graphVAE_train.py:82-88 (500..4000)
    assert [i for _, i in ds.framelist] == [0, 1, 2]
    pc, nor, frame, idx = ds[1]
    mesh = torch.from_numpy(P.read_ply(ds.mesh_path(731))["vertices"])
    want_pc = mesh * 0.001                                                     # :113 SCALE
    ft = torch.from_numpy(f).long()
    v0, v1, v2 = want_pc[ft[:, 0]], want_pc[ft[:, 1]], want_pc[ft[:, 2]]       # :115-120, op for op
    want_nor = (v1 - v0).cross(v2 - v1, dim=1)
    want_nor = want_nor / (want_nor.pow(2).sum(1, keepdim=True).sqrt() + 1e-6)
    assert torch.equal(pc, want_pc) and torch.allclose(nor, want_nor, rtol=0, atol=0) and frame == "731" and idx == 1
    ds2 = P.MeshDataset(track, ["600"], f)
    with pytest.raises(FileNotFoundError):
        ds2[0]


def test_host_batches(tmp_path):
    v, f = _mesh(60)
    track = str(tmp_path) + "/"
    for fr in range(500, 507):
        P.write_ply("%s%06d.ply" % (track, fr), v * (1 + 0.01 * (fr - 500)), f)
    ds = P.MeshDataset(track, list(range(500, 507)), f)
    got = list(P.host_batches(ds, 3, shuffle=True, seed=5, pin=False))
    assert len(got) == 2 and tuple(got[0][0].shape) == (3, 60, 3) and tuple(got[0][1].shape) == (3, len(f), 3)
    again = list(P.host_batches(ds, 3, shuffle=True, seed=5, pin=False))
    assert all(torch.equal(a[0], b[0]) for a, b in zip(got, again))            # same seed, same order
    other = list(P.host_batches(ds, 3, shuffle=True, seed=6, pin=False))
    assert not all(torch.equal(a[0], b[0]) for a, b in zip(got, other))
    full = list(P.host_batches(ds, 3, shuffle=False, drop_last=False, pin=False))
    assert [b[0].shape[0] for b in full] == [3, 3, 1]
    # a torch DataLoader over the same dataset works as in the reference (:295)
    dl = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False)
    pc, nor, frame, idx = next(iter(dl))
    assert tuple(pc.shape) == (2, 60, 3) and tuple(nor.shape) == (2, len(f), 3)
