"""Reader for the reference's INI training config (same sections and keys as
GraphAutoEncoder/graphAE_param_iso.py:21-134 / config_train.config), so a config
written for the reference drives this trainer unchanged.

Differences, all deliberate: no side effects besides creating the two output
folders (the reference also opens a tensorboardX writer here, :33-34); a layer
name that matches no file in `connection_folder` raises instead of printing
"!!!ERROR" and continuing (:112-113); data-set keys that are absent are simply
left unset (vcmeshconv_b200/train.py needs template_ply_fn, mesh_train and
frame_list; tensors can be fed to the trainer directly without them).
"""
import configparser
import json
import os

import numpy as np


class Parameters:
    # (section, key, attribute, type); `None` type = string
    _FIELDS = [
        ("Record", "read_weight_path", "read_weight_path", None),
        ("Record", "write_weight_folder", "write_weight_folder", None),
        ("Record", "write_tmp_folder", "write_tmp_folder", None),
        ("Record", "logdir", "logdir", None),
        ("Params", "lr", "lr", float), ("Params", "batch", "batch", int),
        ("Params", "augment_data", "augmented_data", int),
        ("Params", "start_iter", "start_iter", int), ("Params", "end_iter", "end_iter", int),
        ("Params", "evaluate_iter", "evaluate_iter", int), ("Params", "save_weight_iter", "save_weight_iter", int),
        ("Params", "save_tmp_iter", "save_tmp_iter", int),
        ("Params", "w_pose", "w_pose", float), ("Params", "w_laplace", "w_laplace", float),
        ("Params", "w_color", "w_color", float), ("Params", "w_w_weights_l1", "w_w_weights_l1", float),
        ("Params", "pcs_train", "pcs_train", None), ("Params", "pcs_evaluate", "pcs_evaluate", None),
        ("Params", "pcs_mean", "pcs_mean", None), ("Params", "template_ply_fn", "template_ply_fn", None),
        ("Params", "template_obj_fn", "template_obj_fn", None), ("Params", "mesh_train", "mesh_train", None),
        ("Params", "recon_train", "recon_train", None), ("Params", "calib_path", "calib_path", None),
        ("Params", "frame_list", "framelist", None),
        ("Params", "point_num", "point_num", int), ("Params", "residual_rate", "residual_rate", float),
        ("Params", "conv_max", "conv_max", int), ("Params", "perpoint_bias", "perpoint_bias", int),
        ("Params", "minus_smoothed", "minus_smoothed", int),
    ]
    _REQUIRED = {"lr", "batch", "w_pose", "w_laplace", "point_num", "residual_rate", "conv_max", "perpoint_bias"}

    def read_config(self, fn, make_dirs=True):
        cfg = configparser.ConfigParser()
        if not cfg.read(fn):
            raise FileNotFoundError(fn)
        for sec, key, attr, typ in self._FIELDS:
            if cfg.has_option(sec, key):
                raw = cfg.get(sec, key).strip()
                setattr(self, attr, raw if typ is None else typ(raw))
            elif attr in self._REQUIRED:
                raise KeyError("[%s] %s missing in %s" % (sec, key, fn))
        if make_dirs:
            for d in (getattr(self, "write_weight_folder", ""), getattr(self, "write_tmp_folder", "")):
                if d:
                    os.makedirs(d, exist_ok=True)
        self.freeze_decoder = cfg.getint("Params", "freeze_decoder", fallback=0)
        self.freeze_start_layer = cfg.getint("Params", "freeze_start_layer", fallback=1234567)

        # layer-0 table: graph Laplacian of the losses (graphAE_param_iso.py:88-93); overrides point_num
        self.initial_connection_fn = cfg.get("Params", "initial_connection_fn").strip()
        t0 = np.load(self.initial_connection_fn)
        self.point_num = t0.shape[0]
        self.neighbor_id_lstlst = t0[:, 1:].reshape(self.point_num, -1, 2)[:, :, 0]
        self.neighbor_num_lst = np.array(t0[:, 0])

        self.connection_folder = cfg.get("Params", "connection_folder").strip()
        self.connection_layer_lst_enc = json.loads(cfg.get("Params", "connection_layer_lst_enc"))
        self.connection_layer_lst_dec = json.loads(cfg.get("Params", "connection_layer_lst_dec"))
        files = sorted(os.listdir(self.connection_folder))

        def resolve(names):
            out = []
            for name in names:
                tag = "_" + name + "."
                hit = [f for f in files if tag in f and (".npy" in f or ".npz" in f)]
                if not hit:
                    raise FileNotFoundError("no connection file for layer %r in %s" % (name, self.connection_folder))
                out.append(os.path.join(self.connection_folder, hit[0]))
            return out

        self.connection_layer_fn_lst_enc = resolve(self.connection_layer_lst_enc)
        self.connection_layer_fn_lst_dec = resolve(self.connection_layer_lst_dec)
        for key in ("channel_lst_enc", "channel_lst_dec", "weight_num_lst_enc", "weight_num_lst_dec"):
            if cfg.has_option("Params", key):
                setattr(self, key, json.loads(cfg.get("Params", key)))
        return self
