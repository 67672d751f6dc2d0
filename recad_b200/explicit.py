"""Explicit (rating-matrix) attack dataset: drop-in for recad/dataset/explicit.py `ExplicitData`, the dataset every
trainable attacker iterates (`generate_batch` -> {"users", "users_mat"} row batches of the dense rating matrix).

Same constructor keys, `from_config(name, **kw)`, in-memory entry point (`train_dict` / `valid_dict` / `test_dict` as
float [n, 3] arrays of (user, item, rating), explicit.py:47-52), `.npy` cache names (explicit.py:53-64), `remap_enable`,
`partial_sample(user_ratio=)`, `info_describe()` keys and batch order (np.random.permutation of the filtered users on the
global generator, explicit.py:166-188).

What differs is WHERE a batch's `users_mat` is built.  The reference slices the host matrix, widens it to float64, narrows
it back to a float32 tensor and uploads it for every batch (explicit.py:177-186: ~1.5 ms per 256 x 3702 batch); here the
matrix is uploaded ONCE and a batch's `users_mat` is a device row gather made the first time a consumer asks for the key.
`recad_b200.attacker.Aush` works on the sparse form and never asks, the reference's own attackers get the same tensor as
before.
"""
import os

import numpy as np
import torch

from .config import explicit_defaults, get_logger, merge_config


class _RowBatch(dict):
    """{"users": int64 [B] on the device, "users_mat": float32 [B, n_items]} with `users_mat` made on first use."""

    def __init__(self, users, make):
        super().__init__(users=users)
        self._make = make

    def __missing__(self, key):
        if key != "users_mat":
            raise KeyError(key)
        self["users_mat"] = self._make()
        return dict.__getitem__(self, key)

    def _all(self):
        self["users_mat"]
        return self

    def __contains__(self, key):
        return key in ("users", "users_mat")

    def get(self, key, default=None):
        return self[key] if key in self else default

    def keys(self):
        return dict.keys(self._all())

    def items(self):
        return dict.items(self._all())

    def values(self):
        return dict.values(self._all())

    def __iter__(self):
        return dict.__iter__(self._all())

    def __len__(self):
        return 2


class ExplicitData:
    def __init__(self, path_train, path_test, path_valid, header, sep, threshold, logging_level, **config):
        self.path_train, self.path_test, self.path_valid = path_train, path_test, path_valid
        self.header = header if header is not None else ["user_id", "item_id", "rating"]
        self.sep, self.threshold = sep, threshold
        self.logger = get_logger(f"{__name__}:{self.dataset_name}", level=logging_level)
        self._mode = "train"
        self.config = config
        self.remap_enable = config["remap_enable"]
        self._mat_dev = None
        self._load_data()

    # ---------------------------------------------------------------- factory / reset (dataset/base.py:15-49, 108-118)
    @classmethod
    def from_config(cls, name, **user_config):
        defaults = explicit_defaults(name)
        cfg = merge_config(defaults, {k: v for k, v in user_config.items() if k != "download"}, logger=None)
        inst = object.__new__(cls)
        inst._dataset_name = name
        inst._init_config = cfg
        inst.__init__(**cfg)
        return inst

    @property
    def dataset_name(self):
        return getattr(self, "_dataset_name", type(self).__name__)

    def reset(self, **kwargs):
        config = dict(self._init_config)
        for k, v in kwargs.items():
            if k not in config:
                raise ValueError(f"reset arg {k} should be in {list(config)}")
            config[k] = v
        return type(self).from_config(self.dataset_name, **config)

    # ---------------------------------------------------------------- explicit.py:41-119
    def load_file_as_np(self, file):
        import pandas as pd
        return pd.read_csv(file, engine="python", sep=self.sep).loc[:, self.header].to_numpy()

    def _load_data(self):
        for attr, default in zip(["train_dict", "valid_dict", "test_dict"], [self.path_train, self.path_valid, self.path_test]):
            if self.config[attr] is not None:
                setattr(self, attr, self.config[attr])
                continue
            cache = os.path.join(self.config["cache_dir"], f"{self.dataset_name}_explicit_{attr}.npy")
            if os.path.exists(cache) and self.config["if_cache"]:
                setattr(self, attr, np.load(cache))
            else:
                setattr(self, attr, self.load_file_as_np(default))
                os.makedirs(self.config["cache_dir"], exist_ok=True)
                if self.config["if_cache"]:
                    np.save(cache, getattr(self, attr))
        if self.remap_enable:                          # explicit.py:66-75: user ids -> their rank among all ids seen, IN PLACE
            ids = np.unique(np.concatenate([self.train_dict[:, 0], self.test_dict[:, 0], self.valid_dict[:, 0]]))
            self.user_map = {ids[i]: i for i in range(len(ids))}
            for kvr in (self.train_dict, self.test_dict, self.valid_dict):
                kvr[:, 0] = np.searchsorted(ids, kvr[:, 0])
        self.n_users = int(max(self.train_dict[:, 0].max(), self.valid_dict[:, 0].max(), self.test_dict[:, 0].max()) + 1)
        self.n_items = int(max(self.train_dict[:, 1].max(), self.valid_dict[:, 1].max(), self.test_dict[:, 1].max()) + 1)
        self.train_size, self.valid_size, self.test_size = len(self.train_dict), len(self.valid_dict), len(self.test_dict)
        self.train_mat = self.to_matrix(self.train_dict, self.n_users, self.n_items)

    @staticmethod
    def to_matrix(kv_array, n_users, n_items):
        """explicit.py:110-119: dense float32 [n_users, n_items]; repeated (user, item) rows add up, as scipy's csr does."""
        mat = np.zeros((n_users, n_items), dtype=np.float32)
        np.add.at(mat, (kv_array[:, 0].astype("int64"), kv_array[:, 1].astype("int64")), kv_array[:, 2].astype("float32"))
        return mat

    # ---------------------------------------------------------------- describe
    def batch_describe(self):
        if self.mode() == "train":
            return {"users": (torch.int64, "VarDim(max=batch_size)"),
                    "users_mat": (torch.float32, ("VarDim(max=batch_size)", self.n_items))}

    def info_describe(self):
        infos = {"n_users": self.n_users, "n_items": self.n_items, "train_interactions": self.train_size,
                 "valid_interactions": self.valid_size, "test_interactions": self.test_size, "train_kvr": self.train_dict,
                 "train_mat": self.train_mat, "batch_describe": self.batch_describe()}
        if self.remap_enable:
            infos["user_map"] = self.user_map
        return infos

    def print_help(self, **kwargs):
        from pprint import pprint
        pprint({k: (v if np.isscalar(v) else type(v)) for k, v in self.info_describe().items()})

    def mode(self):
        return self._mode

    def switch_mode(self, mode):
        assert mode in ["train", "test", "validate"]
        self._mode = mode

    # ---------------------------------------------------------------- explicit.py:166-188
    def _rows(self, users):
        dev = torch.device(self.config["device"])
        if dev.type != "cuda":
            return torch.tensor(self.train_mat[users.cpu().numpy(), :].astype("float"), dtype=torch.float32)
        if self._mat_dev is None or self._mat_dev.device != users.device:
            self._mat_dev = torch.from_numpy(self.train_mat).to(users.device)
        return self._mat_dev.index_select(0, users)

    def generate_batch(self, **config):
        user_filter = config.get("user_filter", None)
        if self.mode() == "train":
            batch_size = self.config["batch_size"]
            available_idx = user_filter(train_mat=self.train_mat) if user_filter is not None else list(range(len(self.train_mat)))
            available_idx = np.random.permutation(available_idx)
            users_all = torch.tensor(available_idx, dtype=torch.int64).to(self.config["device"])       # ONE upload per epoch
            for b in range((len(available_idx) + batch_size - 1) // batch_size):
                users = users_all[b * batch_size:(b + 1) * batch_size]
                yield _RowBatch(users, lambda users=users: self._rows(users))

    def inject_data(self, mode, data):
        raise NotImplementedError             # dataset/base.py:82-91 (explicit.py:193-195 defers to it)

    def delete_data(self, mode, user_id, data):
        raise NotImplementedError

    def partial_sample(self, **kwargs):
        """explicit.py:201-230: keep the rows of a random user_ratio share of the train users, remapped."""
        assert "user_ratio" in kwargs, "Expect to have [user_ratio]"
        users = np.unique(self.train_dict[:, 0])
        np.random.shuffle(users)
        left = users[: int(len(users) * kwargs["user_ratio"])]
        keep = lambda kvr: kvr[np.isin(kvr[:, 0], left)]
        return self.reset(train_dict=keep(self.train_dict), test_dict=keep(self.test_dict), valid_dict=keep(self.valid_dict),
                          remap_enable=True, if_cache=False)
