"""Defaults and the lazy-instantiation contract of the reference, restated.

The reference keeps three nested default dicts (recad/default.py:49-89 DATASET,
103-233 MODEL, 247-282 WORKFLOW) and a `from_config` that overlays user kwargs on
a copy of the defaults (recad/model/base.py:27-53, recad/dataset/base.py:15-49).
Only the entries of the victim hot path are mirrored here, with the same key
names and values, so this package works where the reference is not installed
(the GPU box) and plugs into the reference's registries where it is
(recad_b200/register.py).
"""
import logging
import os
from copy import copy

import torch

DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")   # default.py:19
SEED = 2023                                                             # default.py:21


def _dataset_root():
    return os.path.join(os.environ.get("RECAD_DIR", "./"), "data")      # default.py:17-18


def implicit_defaults(name):
    """DATASET['implicit'][name] (default.py:49-63, 82-89) + one new key."""
    root = _dataset_root()
    return {
        "path_train": os.path.abspath(os.path.join(root, name, f"{name}_train.csv")),
        "path_valid": os.path.abspath(os.path.join(root, name, f"{name}_valid.csv")),
        "path_test": os.path.abspath(os.path.join(root, name, f"{name}_test.csv")),
        "test_batch_size": 400,
        "A_split": False,
        "A_n_fold": 100,
        "pairwise_batch_size": 1024,
        "pointwise_batch_size": 1024,
        "sample": "pairwise",
        "negative_ratio": 4,
        "need_graph": True,
        "rating_filter": 4,
        "logging_level": logging.INFO,
        "train_dict": None,
        "valid_dict": None,
        "test_dict": None,
        "remap_enable": False,
        "device": DEVICE,
        "if_cache": False,
        "cache_dir": os.path.abspath(os.path.join(".", "generated")),
        # NEW (not in the reference): which interactions feed the graph / allPos.
        #   "reference": whatever the reference feeds -- the LAST split read, i.e. the test
        #                interactions (implicit.py:233-235 overwrite quirk);
        #   "train":     the intended semantics (train interactions).
        "graph_edges": "reference",
        # NEW: draw the next epochs' samples on background threads while the GPU trains (None = for large datasets)
        "prefetch": None,
    }


def explicit_defaults(name):
    """DATASET['explicit'][name] (default.py:67-89); path_valid is the TEST file there too."""
    root = _dataset_root()
    return {
        "path_train": os.path.abspath(os.path.join(root, name, f"{name}_train.csv")),
        "path_test": os.path.abspath(os.path.join(root, name, f"{name}_test.csv")),
        "path_valid": os.path.abspath(os.path.join(root, name, f"{name}_test.csv")),
        "batch_size": 256, "header": None, "sep": ",", "threshold": 4, "sample": "row",
        "logging_level": logging.INFO, "train_dict": None, "valid_dict": None, "test_dict": None, "remap_enable": False,
        "device": DEVICE, "if_cache": False, "cache_dir": os.path.abspath(os.path.join(".", "generated")),
    }


MODEL = {   # default.py:104-133 (+ logging_level / device, 232-233)
    "victim": {
        "lightgcn": {
            "latent_dim_rec": 128, "lightGCN_n_layers": 3, "A_split": False, "pretrain": False, "keep_prob": 0.6,
            "dropout": 0.0, "lambda": 0.0001, "optim": "adam", "lr": 0.001,
            "init_on_device": False,     # NEW: draw the initial tables on the GPU (not the reference's CPU stream)
        },
        "mf": {"factor_num": 3, "embedding_size": 128, "dropout": 0, "optim": "adam", "lr": 0.001},
        "ncf": {"factor_num": 32, "num_layers": 5, "dropout": 0, "model": "NeuMF-end", "GMF_model": None,
                "MLP_model": None, "optim": "adam", "lr": 0.001,
                # NEW: "tf32x3" = tower GEMMs on the tensor cores (fp32-accurate, ~2^-21); "fp32" = exact CUDA-core
                # GEMMs whose ReLU masks are bit-stable (element-wise parity with the reference's fp32)
                "tower_precision": "tf32x3"},
    },
    "attacker": {   # default.py:159-168
        "aush": {"attack_num": 50, "filler_num": 36, "lr_g": 0.01, "lr_d": 0.001, "optim_g": "adam", "optim_d": "adam",
                 "selected_ids": [62], "ZR_ratio": 0.2},
    },
}
for _scope in MODEL.values():
    for _m in _scope.values():
        _m["logging_level"] = logging.INFO
        _m["device"] = DEVICE

WORKFLOW = {   # default.py:247-282
    "no defense": {
        "rec_epoch": 400, "attack_epoch": 100, "target_id_list": [0], "filter_num": 4, "topks": [10, 20, 50, 100],
        "logging_level": logging.INFO, "device": DEVICE, "verbose": True,
        "validate_every": 0, "validate_k": 20,       # per-epoch Recall/NDCG on the validation split (0 = off, as the reference)
        "cache_dir": os.path.abspath(os.path.join(".", "workflows_results")),
    },
    "defense": {
        "rec_epoch": 400, "attack_epoch": 100, "target_id_list": [0], "filter_num": 4, "topks": [10, 20, 50, 100],
        "defense_epoch": 1, "logging_level": logging.INFO, "device": DEVICE, "verbose": True,
        "validate_every": 0, "validate_k": 20,
        "cache_dir": os.path.abspath(os.path.join(".", "workflows_results")),
    },
}


class NotInstantiatedError(Exception):   # recad/utils.py:19-20
    pass


class InstantiateFail(Exception):        # recad/utils.py:23-24
    pass


def get_logger(name, level=None):
    logger = logging.getLogger(name)
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter("%(asctime)s %(name)s %(levelname)s %(message)s", datefmt="%H:%M:%S"))
        logger.addHandler(handler)
    logger.setLevel(level or logging.INFO)
    return logger


def merge_config(defaults, user_config, user_args=(), logger=None, owner=""):
    """from_config overlay (model/base.py:27-53): copy defaults, apply the user keys that
    are known (defaults or declared user args); unknown keys are logged at debug and dropped."""
    cfg = {k: copy(v) for k, v in defaults.items()}
    for k, v in user_config.items():
        if k in defaults or k in user_args:
            cfg[k] = v
        elif logger is not None:
            logger.debug(f"Unexpected key [{k}] for {owner}")
    return cfg


class LazyMixin:
    """The reference's lazy `.I()` contract (recad/utils.py:200-269) without class patching:
    `from_config(**kw)` returns an un-instantiated shell that remembers its config;
    `.I(**kw)` returns a NEW, fully constructed object (or `self` if already constructed);
    any other public call on a shell raises NotInstantiatedError;
    a constructor failure surfaces as InstantiateFail."""

    _is_instantiate = False

    @classmethod
    def _shell(cls, config, name):
        obj = cls.__new__(cls)
        torch.nn.Module.__init__(obj) if isinstance(obj, torch.nn.Module) else None
        obj._init_config = config
        obj._model_name = name
        obj._is_instantiate = False
        return obj

    def I(self, **kwargs):
        if self._is_instantiate:
            return self
        config = dict(self._init_config)
        config.update(kwargs)
        inst = type(self)._shell(self._init_config, self._model_name)
        try:
            inst._construct(**config)
        except Exception as e:   # utils.py:203-208 wraps every constructor error
            raise InstantiateFail(f"{type(e)}: {e}") from e
        inst._is_instantiate = True
        return inst

    def _require_instance(self, what):
        if not self._is_instantiate:
            raise NotInstantiatedError(f"{type(self).__name__}.{what} is not enabled since no instantiated")
