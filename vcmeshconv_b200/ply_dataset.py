"""Data path of the training script: PLY meshes -> (vertices, face normals) batches.

Mirrors `Dataset` of GraphAutoEncoder/graphVAE_train.py:76-123 and the template loading of :265-268 without the
`plyfile` dependency (absent here, and a 6 890-vertex mesh needs nothing more than a header parser and one
`np.frombuffer`):

  read_ply / write_ply      ascii and binary_little_endian PLY, vertex x/y/z (+ any other scalar properties, skipped)
                            and triangle faces (`vertex_indices` / `vertex_index` list property)
  MeshDataset               frames `{trackpath}{frame:06d}.ply` inside [frame_first, frame_last] (the reference
                            hard-codes 500..4000, :68-73), vertices scaled by SCALE = 0.001 (mm -> m, :30,113), face
                            normals nor = (v1-v0) x (v2-v1) / (|.| + 1e-6) (:115-120); a missing file raises (the
                            reference prints and then fails on an unbound name)
  host_batches              shuffled, pinned host batches for DataParallelTrainer.prefetch / step_prefetched
                            (replaces the 20-worker DataLoader of :295: decoding a frame is one frombuffer)

Everything here runs on the host; the device work starts at DataParallelTrainer.prefetch.
"""
import os

import numpy as np
import torch

SCALE = 0.001                                    # graphVAE_train.py:30

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
              "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def _parse_header(f):
    if f.readline().strip() != b"ply":
        raise ValueError("not a PLY file")
    fmt, elements = None, []
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PLY header is not terminated")
        tok = line.decode("ascii", "replace").split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
        elif tok[0] == "property":
            if not elements:
                raise ValueError("PLY property before any element")
            if tok[1] == "list":
                elements[-1]["props"].append(("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]], tok[4]))
            else:
                elements[-1]["props"].append(("scalar", _PLY_TYPES[tok[1]], None, tok[2]))
        elif tok[0] == "end_header":
            break
    if fmt not in ("ascii", "binary_little_endian"):
        raise ValueError("unsupported PLY format %r (ascii and binary_little_endian are)" % (fmt,))
    return fmt, elements


def _read_scalar_element(f, fmt, el):
    names = [p[3] for p in el["props"]]
    if fmt == "ascii":
        rows = [f.readline().split() for _ in range(el["count"])]
        arr = np.array(rows, dtype=np.float64).reshape(el["count"], len(names))
        return {n: arr[:, i] for i, n in enumerate(names)}
    dt = np.dtype([(n, "<" + p[1]) for n, p in zip(names, el["props"])])
    raw = f.read(dt.itemsize * el["count"])
    if len(raw) != dt.itemsize * el["count"]:
        raise ValueError("PLY file is truncated in element %r" % el["name"])
    rec = np.frombuffer(raw, dtype=dt)
    return {n: rec[n] for n in names}


def _read_face_element(f, fmt, el):
    lists = [p for p in el["props"] if p[0] == "list"]
    if len(lists) != 1 or lists[0][3] not in ("vertex_indices", "vertex_index"):
        raise ValueError("face element needs one list property vertex_indices")
    if fmt == "ascii":
        faces = np.empty((el["count"], 3), np.int64)
        for i in range(el["count"]):
            tok = f.readline().split()
            if int(tok[0]) != 3:
                raise ValueError("face %d is not a triangle" % i)
            faces[i] = [int(t) for t in tok[1:4]]
        return faces
    if any(p[0] != "list" for p in el["props"]):
        raise ValueError("binary face elements with extra scalar properties are not supported")
    _, ct, it, _ = lists[0]
    dt = np.dtype([("n", "<" + ct), ("v", "<" + it, (3,))])           # all faces must be triangles
    raw = f.read(dt.itemsize * el["count"])
    if len(raw) != dt.itemsize * el["count"]:
        raise ValueError("PLY file is truncated in the face element")
    rec = np.frombuffer(raw, dtype=dt)
    if el["count"] and not np.all(rec["n"] == 3):
        raise ValueError("only triangle meshes are supported")
    return rec["v"].astype(np.int64)


def read_ply(path):
    """-> {"vertices": float32 [V,3], "faces": int64 [F,3] or None}."""
    with open(path, "rb") as f:
        fmt, elements = _parse_header(f)
        out = {"vertices": None, "faces": None}
        for el in elements:
            if el["name"] == "vertex":
                if any(p[0] == "list" for p in el["props"]):
                    raise ValueError("list properties on vertices are not supported")
                cols = _read_scalar_element(f, fmt, el)
                if not all(k in cols for k in "xyz"):
                    raise ValueError("vertex element has no x/y/z")
                out["vertices"] = np.stack([cols["x"], cols["y"], cols["z"]], 1).astype(np.float32)
            elif el["name"] == "face":
                out["faces"] = _read_face_element(f, fmt, el)
            elif el["count"]:
                if any(p[0] == "list" for p in el["props"]):
                    raise ValueError("cannot skip list element %r" % el["name"])
                _read_scalar_element(f, fmt, el)
    if out["vertices"] is None:
        raise ValueError("%s has no vertex element" % path)
    return out


def write_ply(path, vertices, faces=None, binary=True):
    """Writes float32 vertices [V,3] (+ int32 triangles [F,3])."""
    v = np.ascontiguousarray(vertices, dtype="<f4").reshape(-1, 3)
    fc = None if faces is None else np.ascontiguousarray(faces, dtype="<i4").reshape(-1, 3)
    hdr = ["ply", "format %s 1.0" % ("binary_little_endian" if binary else "ascii"), "element vertex %d" % len(v),
           "property float x", "property float y", "property float z"]
    if fc is not None:
        hdr += ["element face %d" % len(fc), "property list uchar int vertex_indices"]
    hdr.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        if binary:
            f.write(v.tobytes())
            if fc is not None:
                rec = np.empty(len(fc), dtype=[("n", "u1"), ("v", "<i4", (3,))])
                rec["n"], rec["v"] = 3, fc
                f.write(rec.tobytes())
        else:
            for row in v:
                f.write(("%.9g %.9g %.9g\n" % tuple(row)).encode("ascii"))
            if fc is not None:
                for row in fc:
                    f.write(("3 %d %d %d\n" % tuple(row)).encode("ascii"))


def load_template_faces(template_ply_fn):
    """graphVAE_train.py:265-268: the triangle list of the template mesh as an int64 tensor [F,3]."""
    faces = read_ply(template_ply_fn)["faces"]
    if faces is None:
        raise ValueError("template %s has no faces" % template_ply_fn)
    return torch.from_numpy(faces)


def face_normals(pc, faces):
    """graphVAE_train.py:115-120 on one mesh [V,3] (torch): nor = (v1-v0) x (v2-v1), normalised with +1e-6."""
    v0, v1, v2 = pc[faces[:, 0]], pc[faces[:, 1]], pc[faces[:, 2]]
    nor = torch.linalg.cross(v1 - v0, v2 - v1)
    return nor / (nor.pow(2).sum(1, keepdim=True).sqrt() + 1e-6)


class MeshDataset(torch.utils.data.Dataset):
    """`Dataset` of graphVAE_train.py:76-123. Items: (pc [V,3] in metres, face normals [F,3], frame, index)."""

    def __init__(self, trackpath, framelist, faces, frame_first=500, frame_last=4000, scale=SCALE):
        self.trackpath = trackpath
        self.scale = scale
        self.framelist = []
        for x in framelist:
            if frame_first <= int(x) <= frame_last:                   # :82-88
                self.framelist.append((x, len(self.framelist)))
        self.faceidx = torch.as_tensor(np.asarray(faces)).long()

    def __len__(self):
        return len(self.framelist)

    def mesh_path(self, frame):
        return "%s%06d.ply" % (self.trackpath, int(frame))

    def __getitem__(self, idx):
        frame, indexfr = self.framelist[idx]
        name = self.mesh_path(frame)
        if not os.path.isfile(name):
            raise FileNotFoundError("cannot find %s" % name)
        pc = torch.from_numpy(read_ply(name)["vertices"]) * self.scale
        return pc, face_normals(pc, self.faceidx), frame, indexfr


def host_batches(dataset, batch, shuffle=True, seed=0, drop_last=True, pin=None):
    """One epoch of (pc [B,V,3], nor [B,F,3]) host batches in pinned memory (when CUDA is there), ready for
    DataParallelTrainer.prefetch(); `seed` + the epoch number should be passed by the caller for a new order."""
    pin = torch.cuda.is_available() if pin is None else pin
    order = np.arange(len(dataset))
    if shuffle:
        np.random.default_rng(seed).shuffle(order)
    for s in range(0, len(order), batch):
        ids = order[s:s + batch]
        if len(ids) < batch and drop_last:
            return
        items = [dataset[int(i)] for i in ids]
        pc = torch.stack([it[0] for it in items]).contiguous()
        nor = torch.stack([it[1] for it in items]).contiguous()
        yield (pc.pin_memory(), nor.pin_memory()) if pin else (pc, nor)
