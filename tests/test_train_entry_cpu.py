"""Host side of the training entry point (vcmeshconv_b200/train.py): a reference-format config + PLY frames give
the model, the dataset and per-rank batches; without a GPU main() refuses to run. CPU only."""
import numpy as np
import pytest
import torch

from test_config_cpu import CONFIG
from vcmeshconv_b200 import graph_sampling as gs
from vcmeshconv_b200 import ply_dataset, synthetic, train


def _project(tmp_path, n_vert=400):
    pts, faces = synthetic.sphere(n_vert, seed=0)
    tabs, nums = gs.build_tables(faces, n_vert, synthetic.alternating_levels(4))
    gs.save_tables(tabs, str(tmp_path / "connections" / "body") + "/")
    (tmp_path / "track").mkdir()
    ply_dataset.write_ply(str(tmp_path / "template.ply"), pts * 1000, faces)
    rng = np.random.default_rng(1)
    frames = list(range(498, 506))
    for fr in frames:
        ply_dataset.write_ply("%s/track/final_%06d.ply" % (tmp_path, fr),
                              (pts + 0.01 * rng.standard_normal(pts.shape)).astype(np.float32) * 1000, faces)
    (tmp_path / "frames.txt").write_text("\n".join(str(f) for f in frames) + "\n")
    cfg = CONFIG.format(root=str(tmp_path))
    cfg = cfg.replace("template_ply_fn: testMC/hr_000500.ply", "template_ply_fn: %s/template.ply" % tmp_path)
    cfg = cfg.replace("mesh_train: track/final_", "mesh_train: %s/track/final_" % tmp_path)
    cfg = cfg.replace("frame_list: testMC/filterfrmlistfull.txt", "frame_list: %s/frames.txt" % tmp_path)
    fn = tmp_path / "config_train.config"
    fn.write_text(cfg)
    return str(fn), pts, faces, nums


def test_build_from_config(tmp_path):
    fn, pts, faces, nums = _project(tmp_path)
    param, model, dataset = train.build_from_config(fn, enc_channels=[3, 16, 8], dec_channels=[8, 16, 3], weight_num=5)
    assert param.point_num == 400 and param.batch == 2 and abs(param.lr - 1e-4) < 1e-12
    assert len(dataset) == 6                                  # frames 500..505 of 498..505 (window 500..4000)
    assert torch.equal(dataset.faceidx, torch.from_numpy(faces.astype(np.int64)))
    pc, nor, frame, idx = dataset[0]
    assert tuple(pc.shape) == (400, 3) and tuple(nor.shape) == (len(faces), 3) and frame == "500" and idx == 0
    assert abs(float(pc.norm(dim=1).mean()) - 1.0) < 0.05     # mm in the files, metres in the batch (unit sphere)
    # the model has the reference's sub-modules, on the config's tables
    assert model.net_geoenc.layer_num == 2 and model.net_geodec.layer_num == 2
    assert model.mcstructureenc.ptnum_list[0] == 400 and model.net_loss.initial_neighbor_id_lstlst.shape[0] == 400
    assert train.read_frame_list(param.framelist) == [str(f) for f in range(498, 506)]


def test_rank_slices_partition_the_global_batch(tmp_path):
    fn, *_ = _project(tmp_path)
    _, _, dataset = train.build_from_config(fn, enc_channels=[3, 16, 8], dec_channels=[8, 16, 3], weight_num=5)
    world, per = 2, 2
    got = list(ply_dataset.host_batches(dataset, per * world, shuffle=True, seed=3, pin=False))
    assert len(got) == 1                                      # 6 frames, global batch 4, the remainder is dropped
    pc, nor = got[0]
    parts = [train.rank_slice(pc, nor, r, world) for r in range(world)]
    assert all(p[0].shape[0] == per and p[1].shape[0] == per for p in parts)
    assert torch.equal(torch.cat([p[0] for p in parts]), pc) and torch.equal(torch.cat([p[1] for p in parts]), nor)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a GPU")
def test_main_refuses_without_cuda(tmp_path):
    fn, *_ = _project(tmp_path)
    with pytest.raises(SystemExit) as e:
        train.main([fn])
    assert "no CPU path" in str(e.value)
    with pytest.raises(SystemExit):
        train.main([])
