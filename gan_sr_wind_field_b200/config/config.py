"""Reads the reference's ``config/*.ini`` / ``pretrained_models/*/config.ini`` files unchanged and exposes the
same attribute surface as the reference's ``config/config.py`` (``cfg.scale``, ``cfg.generator.num_features``,
``cfg.training.pixel_loss_weight`` …, SURVEY §5 "Config / flags").

Table-driven: each section lists (key, type); missing keys become ``None`` exactly like
``ConfigParser.get*`` with ``allow_no_value=True`` does in the reference (e.g. ``[GENERATOR] conv_mode``).
Unlike the reference, section objects are per-``Config`` instances, not class-level singletons.
"""
from __future__ import annotations

import ast
from configparser import ConfigParser

_B, _I, _F, _S, _L = "bool", "int", "float", "str", "intlist"

_SCHEMA = {
    "GAN": ("gan_config", [
        ("include_pressure", _B), ("include_z_channel", _B), ("include_above_ground_channel", _B),
        ("number_of_z_layers", _I), ("conv_mode", _S), ("start_date", _L), ("end_date", _L),
        ("interpolate_z", _B), ("use_D_feature_extractor_cost", _B), ("enable_slicing", _B), ("slice_size", _I)]),
    "ENV": ("env", [
        ("root_path", _S), ("log_subpath", _S), ("tensorboard_subpath", _S), ("runs_subpath", _S),
        ("generator_load_path", _S), ("discriminator_load_path", _S), ("state_load_path", _S),
        ("fixed_seed", _I)]),
    "GENERATOR": ("generator", [
        ("norm_type", _S), ("act_type", _S), ("layer_mode", _S), ("num_features", _I), ("num_RRDB", _I),
        ("num_RDB_convs", _I), ("RDB_res_scaling", _F), ("RRDB_res_scaling", _F), ("in_num_ch", _I),
        ("out_num_ch", _I), ("RDB_growth_chan", _I), ("hr_kern_size", _I), ("weight_init_scale", _F),
        ("lff_kern_size", _I), ("conv_mode", _S), ("use_mixed_precision", _B),
        ("terrain_number_of_features", _I), ("dropout_probability", _F), ("max_norm", _F)]),
    "DISCRIMINATOR": ("discriminator", [
        ("norm_type", _S), ("act_type", _S), ("layer_mode", _S), ("num_features", _I), ("in_num_ch", _I),
        ("feat_kern_size", _I), ("weight_init_scale", _F), ("conv_mode", _S), ("use_mixed_precision", _B),
        ("dropout_probability", _F)]),
    "TRAINING": ("training", [
        ("resume_training_from_save", _B), ("learning_rate_g", _F), ("learning_rate_d", _F),
        ("adam_weight_decay_g", _F), ("adam_weight_decay_d", _F), ("adam_beta1_g", _F), ("adam_beta1_d", _F),
        ("multistep_lr", _B), ("multistep_lr_steps", _L), ("lr_gamma", _F), ("gan_type", _S),
        ("adversarial_loss_weight", _F), ("d_g_train_ratio", _I), ("d_g_train_period", _I),
        ("pixel_criterion", _S), ("pixel_loss_weight", _F), ("gradient_xy_loss_weight", _F),
        ("gradient_z_loss_weight", _F), ("divergence_loss_weight", _F), ("xy_divergence_loss_weight", _F),
        ("feature_D_loss_weight", _F), ("use_noisy_labels", _B), ("use_one_sided_label_smoothing", _B),
        ("use_instance_noise", _B), ("flip_labels", _B), ("niter", _I), ("val_period", _I),
        ("save_model_period", _I), ("log_period", _I), ("conv_mode", _S), ("train_eval_test_ratio", _F),
        ("feature_D_update_period", _I)]),
}
_DATASET_KEYS = [("name", _S), ("mode", _S), ("dataroot_hr", _S), ("dataroot_lr", _S), ("num_workers", _I),
                 ("batch_size", _I), ("data_aug_flip", _B), ("data_aug_rot", _B)]
_DATASETS = {"DATASETTRAIN": "dataset_train", "DATASETTEST": "dataset_test", "DATASETVAL": "dataset_val"}


def safe_list_from_string(text, target_type=int) -> list:
    """'[1, 2]' -> [1, 2]; '3' -> [3]; anything unparsable / None -> []."""
    try:
        value = ast.literal_eval(text)
    except Exception:
        return []
    if value is None:
        return []
    return list(value) if isinstance(value, (list, tuple)) else [value]


def _read(section, key, kind):
    if kind == _B:
        return section.getboolean(key)
    if kind == _I:
        return section.getint(key)
    if kind == _F:
        return section.getfloat(key)
    if kind == _L:
        return safe_list_from_string(section.get(key), int)
    return section.get(key)


class Section:
    """One ini section as attributes; ``str()`` renders it back as ini text."""

    def __init__(self, title, section, keys):
        self._title = title
        for key, kind in keys:
            setattr(self, key, _read(section, key, kind))

    def __str__(self):
        lines = [f"[{self._title}]"]
        for k, v in vars(self).items():
            if k.startswith("_"):
                continue
            lines.append(f"{k} = {v}" if v is not None else f"{k}")
        return "\n".join(lines) + "\n"


class Config:
    is_train = False
    is_use = False
    is_test = False
    is_param_search = False
    is_download = False
    slurm_array_id = 1
    device = None

    def __init__(self, ini_path):
        parser = ConfigParser(allow_no_value=True)
        if not parser.read(ini_path):
            raise FileNotFoundError(ini_path)
        base = parser["DEFAULT"]
        self.name = base.get("name")
        self.model = base.get("model")
        self.use_tensorboard_logger = base.getboolean("use_tensorboard_logger")
        self.scale = base.getint("scale")
        self.also_log_to_terminal = base.getboolean("also_log_to_terminal")
        gpu = base.get("gpu_id")
        self.gpu_id = None if gpu is None or gpu.lower() == "none" else int(gpu)
        self.load_model_from_save = base.getboolean("load_model_from_save")
        self.display_bar = base.getboolean("display_bar")
        for title, (attr, keys) in _SCHEMA.items():
            setattr(self, attr, Section(title, parser[title], keys))
        for title, attr in _DATASETS.items():
            setattr(self, attr, Section(title, parser[title], _DATASET_KEYS) if parser.has_section(title) else None)

    def asINI(self) -> str:
        return str(self)

    def __str__(self):
        head = ["[DEFAULT]"] + [f"{k} = {v}" for k, v in vars(self).items() if not isinstance(v, Section)
                                and v is not None or k == "gpu_id"]
        parts = ["\n".join(head) + "\n"]
        for attr in ("env", "gan_config", "generator", "discriminator", "training", "dataset_train",
                     "dataset_val", "dataset_test"):
            sec = getattr(self, attr)
            if sec is not None:
                parts.append(str(sec))
        return "\n".join(parts)


# typed aliases so ``config.GeneratorConfig`` etc. resolve for annotations written against the reference
GANConfig = EnvConfig = GeneratorConfig = DiscriminatorConfig = TrainingConfig = DatasetConfig = Section
