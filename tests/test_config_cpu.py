"""The INI reader accepts the reference's config format (config_train.config keys). CPU only."""
import os

import numpy as np
import pytest

from vcmeshconv_b200 import graph_sampling as gs
from vcmeshconv_b200 import synthetic
from vcmeshconv_b200.graphAE_param_iso import Parameters

CONFIG = """[Record]
read_weight_path:

write_weight_folder: {root}/weight_00/
write_tmp_folder: {root}/tmp_00/
logdir: {root}/log_00/

[Params]
lr: 0.0001
batch: 2
w_pose: 1
w_laplace: 100
w_color: 0.0
w_w_weights_l1: 0.0
augment_data: 0
start_iter: 0
end_iter: 1000001
save_weight_iter: 5000
save_tmp_iter: 1000
evaluate_iter: 1000
residual_rate: 0.9
conv_max: 0
perpoint_bias: 1
minus_smoothed: 0
template_ply_fn: testMC/hr_000500.ply
template_obj_fn: testMC/cloth000500.obj
mesh_train: track/final_
recon_train: meshes/plys/
calib_path: testMC/calib.txt
frame_list: testMC/filterfrmlistfull.txt
point_num: 139337
pcs_train: testMC/train_with_weights.npy
pcs_evaluate: testMC/eval_with_weights.npy
pcs_mean:
connection_folder:  {root}/connections/body/
connection_layer_lst_enc: ["pool1", "pool3"]
channel_lst_enc:          [32 ,      64]
weight_num_lst_enc:       [17,      17]
connection_layer_lst_dec: ["unpool3", "unpool1"]
channel_lst_dec:          [32,        6]
weight_num_lst_dec:       [17,       17]
initial_connection_fn: {root}/connections/body/_pool0.npy
"""


def test_reference_style_config(tmp_path):
    pts, faces = synthetic.sphere(400, seed=0)
    tabs, nums = gs.build_tables(faces, 400, synthetic.alternating_levels(4))
    folder = str(tmp_path / "connections" / "body") + "/"
    gs.save_tables(tabs, folder)
    cfg = tmp_path / "train.config"
    cfg.write_text(CONFIG.format(root=str(tmp_path)))
    p = Parameters().read_config(str(cfg))
    assert p.lr == 1e-4 and p.batch == 2 and p.residual_rate == 0.9 and p.perpoint_bias == 1 and p.conv_max == 0
    assert p.point_num == 400                                   # taken from the layer-0 table, as in the reference
    assert np.array_equal(p.neighbor_id_lstlst, tabs["pool0"][:, 1::2]) and np.array_equal(p.neighbor_num_lst, tabs["pool0"][:, 0])
    assert [os.path.basename(f) for f in p.connection_layer_fn_lst_enc] == ["_pool1.npy", "_pool3.npy"]
    assert [os.path.basename(f) for f in p.connection_layer_fn_lst_dec] == ["_unpool3.npy", "_unpool1.npy"]
    assert p.channel_lst_dec == [32, 6] and p.weight_num_lst_enc == [17, 17]
    assert os.path.isdir(p.write_weight_folder) and os.path.isdir(p.write_tmp_folder)
    # the resolved tables feed MCStructure as in the reference
    from vcmeshconv_b200 import graphVAESSW as vae
    s = vae.MCStructure(p, p.point_num, 17, bDec=False)
    assert s.ptnum_list == [400, nums[2], nums[4]]


def test_missing_layer_file_raises(tmp_path):
    pts, faces = synthetic.sphere(200, seed=0)
    tabs, _ = gs.build_tables(faces, 200, synthetic.alternating_levels(2))
    gs.save_tables(tabs, str(tmp_path / "connections" / "body") + "/")
    cfg = tmp_path / "train.config"
    cfg.write_text(CONFIG.format(root=str(tmp_path)))
    with pytest.raises(FileNotFoundError):
        Parameters().read_config(str(cfg))                      # pool3 / unpool3 do not exist
