"""AUSH attacker on the CUDA path: drop-in for recad/model/attacker/aush.py `Aush` (same config keys, `from_config` /
lazy `.I(dataset=...)`, `train_step(target_id_list=...)` returning the same 4-tuple, `generate_fake(target_id_list=...)`
returning the same float32 [attack_num, n_items] array, `input_describe` / `output_describe`).

Division of work per `train_step` (= one pass over the attack dataset's batches, aush.py:100-170):
  host   the draws on the global numpy generator, bit-exact and in the reference's order -- the dataset's own batch
         generator (np.random.permutation, explicit.py:166-188) is consumed as is; per batch `sample_fillers`
         (np.random.choice per row) and the ZR-pool shuffle run in C (recad_mt19937_aush_batch) -- and the per-batch
         column index of the sparse inputs (recad_aush_plan_columns);
  device everything numerical, for the whole epoch in one C call (recad_aush_train_epoch, csrc/aush.cu).
The reference builds four dense [B, n_items] masks per batch on the host and uploads them; here a batch is
B x (filler_num + |selected|) (column, value) pairs.

The generator never trains in the reference (its output is detached before every loss, aush.py:126-128, so
G_optimizer.step() has nothing to apply; golden: tests/golden/make_golden_aush.py) -- it is evaluated, not updated, here too.
"""
import ctypes as C
from functools import partial

import numpy as np
import torch

from . import _lib, ops
from .config import MODEL, LazyMixin, get_logger, merge_config

_G_KEYS = ("main.0.weight", "main.0.bias", "main.2.weight", "main.2.bias")
_D_KEYS = tuple(f"main.{2 * l}.{p}" for l in range(4) for p in ("weight", "bias"))
_H, _HP = 150, 152


def filler_filter_mat(train_mat, target_id_list=(), selected_ids=(), filler_num=0):
    """utils.py:192-196: the rows with at least filler_num rated items outside selected_ids + target_id_list."""
    skip = np.unique(np.asarray(list(selected_ids) + list(target_id_list), dtype=np.int64))
    counts = np.count_nonzero(train_mat > 0, axis=1) - (np.count_nonzero(train_mat[:, skip] > 0, axis=1) if len(skip) else 0)
    return np.where(counts >= filler_num)[0]


def _f32(x):
    """A host float32 tensor from a tensor on any device or anything numpy understands."""
    return x.detach().to("cpu", torch.float32) if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float32))


def _vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Aush(LazyMixin, torch.nn.Module):
    name = "aush"
    user_args = ("dataset",)

    @classmethod
    def from_config(cls, **kwargs):
        """model.from_config('attacker', 'aush', **kw) (aush.py:38-42): lazy shell; `.I(dataset=...)` builds."""
        cfg = merge_config(MODEL["attacker"]["aush"], kwargs, cls.user_args, get_logger(__name__), owner=str(cls))
        return cls._shell(cfg, cls.name)

    @property
    def model_name(self):
        return getattr(self, "_model_name", type(self).__name__)

    def reset(self, **kwargs):
        """model/base.py:94-104."""
        config = dict(self._init_config)
        for k, v in kwargs.items():
            if k not in config:
                raise ValueError(f"reset arg {k} should be in {list(config)}")
            config[k] = v
        return type(self).from_config(**config)

    # ------------------------------------------------------------------ construction (aush.py:12-36)
    def _construct(self, **config):
        torch.nn.Module.__init__(self)
        self.config = config
        self.selected_ids = list(config["selected_ids"])
        self.attack_num, self.filler_num = int(config["attack_num"]), int(config["filler_num"])
        self.dataset, self.ZR_ratio = config["dataset"], config["ZR_ratio"]
        dev = torch.device(config["device"])
        if dev.type != "cuda":
            raise ops.RecadError(f"recad_b200 Aush computes on a CUDA device only (config device = {dev}); there is no CPU fallback")
        self.device = dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        for key in ("optim_g", "optim_d"):
            if str(config[key]).lower() != "adam":
                raise ValueError(f"{key}={config[key]!r}: only 'adam' (the default, default.py:164-165) is implemented")
        info = self.dataset.info_describe()
        self.n_items = int(info["n_items"])
        self.train_data_array = info["train_mat"]
        self._mat = np.ascontiguousarray(self.train_data_array, dtype=np.float32)
        self._sel = np.unique(np.asarray(self.selected_ids, dtype=np.int64))             # slot order: ascending column
        if len(self._sel) > 64 or (len(self._sel) and (self._sel[0] < 0 or self._sel[-1] >= self.n_items)):
            raise ValueError("selected_ids: at most 64 distinct item ids inside [0, n_items)")
        self._step = 0
        self._cand = {}
        self._elig = None
        self._build_network()

    def _build_network(self):
        """aush.py:26-36 + 254-283: both stacks are drawn on the CPU from the global torch generator in the reference's
        order (generator first), then moved."""
        nn, I = torch.nn, self.n_items
        G = nn.Sequential(nn.Linear(I, 128), nn.Sigmoid(), nn.Linear(128, I), nn.Sigmoid())
        D = nn.Sequential(nn.Linear(I, _H), nn.Sigmoid(), nn.Linear(_H, _H), nn.Sigmoid(), nn.Linear(_H, _H), nn.Sigmoid(),
                          nn.Linear(_H, 1), nn.Sigmoid())
        off = (C.c_int64 * 9)()
        _lib.check(_lib.lib().recad_aush_d_layout(I, off), "recad_aush_d_layout")
        self._off = list(off)
        with torch.cuda.device(self.device):
            self._D = torch.zeros(self._off[8], dtype=torch.float32, device=self.device)
            self._Dm, self._Dv = torch.zeros_like(self._D), torch.zeros_like(self._D)
            self._G = [torch.empty(0, device=self.device)] * 4
            self._sel_dev = torch.as_tensor(self._sel.astype(np.int32)).to(self.device)
        self.load_netG_state({f"main.{k}": v.detach() for k, v in G.state_dict().items()})
        self.load_netD_state({f"main.{k}": v.detach() for k, v in D.state_dict().items()})

    # ------------------------------------------------------------------ parameters in the reference's naming / layout
    def load_netG_state(self, sd):
        w1, b1, w2, b2 = (_f32(sd[k]) for k in _G_KEYS)
        self._G = [w1.t().contiguous().to(self.device), b1.contiguous().to(self.device), w2.contiguous().to(self.device),
                   b2.contiguous().to(self.device)]

    def netG_state(self):
        w1t, b1, w2, b2 = self._G
        return dict(zip(_G_KEYS, (w1t.t().contiguous(), b1.clone(), w2.clone(), b2.clone())))

    def _d_views(self, buf):
        o, I = self._off, self.n_items
        return {
            "main.0.weight": buf[o[0]:o[0] + I * _HP].view(I, _HP)[:, :_H],          # stored transposed
            "main.0.bias": buf[o[1]:o[1] + _H],
            "main.2.weight": buf[o[2]:o[2] + _H * _HP].view(_H, _HP)[:, :_H],
            "main.2.bias": buf[o[3]:o[3] + _H],
            "main.4.weight": buf[o[4]:o[4] + _H * _HP].view(_H, _HP)[:, :_H],
            "main.4.bias": buf[o[5]:o[5] + _H],
            "main.6.weight": buf[o[6]:o[6] + _H].view(1, _H),
            "main.6.bias": buf[o[7]:o[7] + 1],
        }

    def load_netD_state(self, sd):
        self._D.zero_()
        v = self._d_views(self._D)
        for k in _D_KEYS:
            t = _f32(sd[k]).to(self.device)
            v[k].copy_(t.t() if k == "main.0.weight" else t)

    def netD_state(self):
        v = self._d_views(self._D)
        return {k: (v[k].t() if k == "main.0.weight" else v[k]).contiguous().clone() for k in _D_KEYS}

    def to(self, device=None, *args, **kwargs):
        self._require_instance("to")
        return self

    # ------------------------------------------------------------------ describe (aush.py:44-57)
    def info_describe(self):
        return {"input_describe": self.input_describe(), "output_describe": self.output_describe()}

    def input_describe(self):
        return {"train_step": {"target_id_list": (list, "VarDim")}}

    def output_describe(self):
        return {"train_step": {"d_losses": (float, []), "g_loss_rec_l": (float, []), "g_loss_shilling_l": (float, []),
                               "g_loss_gan_l": (float, [])}}

    def forward(self):
        pass

    # ------------------------------------------------------------------ host side
    def _eligible(self, train_mat, targets):
        key = (id(train_mat), tuple(targets))
        if self._elig is None or self._elig[0] != key:
            self._elig = (key, filler_filter_mat(train_mat, list(targets), self.selected_ids, self.filler_num), train_mat)
        return self._elig[1]

    def _candidates(self, targets):
        """Per user the list sample_fillers draws from, in the reference's order: list(set(rated columns) & filler_pool)
        (aush.py:61-69) is a CPython set iteration order, so the same expression produces it -- ONCE per user and target
        list instead of once per row and epoch."""
        key = tuple(targets)
        if key not in self._cand:
            pool = set(range(self.n_items)) - set(self.selected_ids) - set(targets)
            ptr = np.zeros(self._mat.shape[0] + 1, dtype=np.int64)
            items, vals = [np.zeros(0, dtype=np.int32)], [np.zeros(0, dtype=np.float32)]
            for u in filler_filter_mat(self._mat, list(targets), self.selected_ids, 1):
                lst = np.asarray(list(set(np.argwhere(self._mat[u] > 0).flatten()) & pool), dtype=np.int32)
                items.append(lst)
                vals.append(self._mat[u, lst])
                ptr[u + 1] = len(lst)
            np.cumsum(ptr, out=ptr)
            self._cand = {key: (ptr, np.concatenate(items), np.concatenate(vals))}
        return self._cand[key]

    def _draw_batch(self, users, targets, with_zr=True):
        """One batch's draws on np.random's global state (recad_mt19937_aush_batch): the filler columns, input_template at
        those columns (the rating, once per distinct column of a row) and the ZR mask at the selected columns."""
        ptr, items, vals = self._candidates(targets)
        B, F, S = len(users), self.filler_num, len(self._sel) if with_zr else 0
        users = np.ascontiguousarray(users, dtype=np.int64)
        cols = np.empty((B, F), dtype=np.int32)
        tval = np.empty((B, F), dtype=np.float32)
        zr = np.empty((B, S), dtype=np.float32)
        zero_sel = np.ascontiguousarray(self._mat[users[:, None], self._sel[None, :]] == 0, dtype=np.uint8) if S else None
        st = np.random.get_state()
        key, pos = np.ascontiguousarray(st[1], dtype=np.uint32).copy(), C.c_int32(int(st[2]))
        _lib.check(_lib.lib().recad_mt19937_aush_batch(_vp(key), C.byref(pos), B, _vp(users), _vp(ptr), _vp(items), _vp(vals), F, S,
                                                       _vp(zero_sel), float(self.ZR_ratio), _vp(cols), _vp(tval), _vp(zr) if S else None),
                   "recad_mt19937_aush_batch")
        np.random.set_state(("MT19937", key, pos.value, 0, 0.0))
        return cols, tval, zr

    def _state(self, work=None):
        st = _lib.Aush()
        st.n_items, st.n_sel, st.filler_num = self.n_items, len(self._sel), self.filler_num
        st.lr, st.beta1, st.beta2, st.eps = float(self.config["lr_d"]), 0.9, 0.999, 1e-8
        st.G_W1t, st.G_b1, st.G_W2, st.G_b2 = (t.data_ptr() for t in self._G)
        st.D, st.Dm, st.Dv = self._D.data_ptr(), self._Dm.data_ptr(), self._Dv.data_ptr()
        st.selected = self._sel_dev.data_ptr()
        st.work = work.data_ptr() if work is not None else None
        return st

    # ------------------------------------------------------------------ aush.py:78-180
    def train_step(self, **config):
        self._require_instance("train_step")
        targets = list(config["target_id_list"])
        index_filter = partial(self._user_filter, targets=targets)
        users_l, cols_l, tval_l, zr_l, batch = [], [], [], [], 0
        for dp in self.dataset.generate_batch(user_filter=index_filter, **config):
            users = dp["users"].cpu().numpy().astype(np.int64)
            cols, tval, zr = self._draw_batch(users, targets)
            users_l.append(users); cols_l.append(cols); tval_l.append(tval); zr_l.append(zr)
            batch = max(batch, len(users))
        if not users_l:
            nan = float("nan")
            return (nan, nan, nan, nan)                          # np.mean([]) of the reference
        if any(len(u) != batch for u in users_l[:-1]):
            raise ops.RecadError("Aush.train_step: the dataset's batches must have one size (the last may be shorter)")
        users, cols, tval, zr = np.concatenate(users_l), np.concatenate(cols_l), np.concatenate(tval_l), np.concatenate(zr_l)
        n, F, S, I = len(users), self.filler_num, len(self._sel), self.n_items
        sel_hit = np.isin(cols, self._sel)
        dval = np.where(sel_hit, np.float32(0), tval)
        fm_sel = (cols[:, :, None] == self._sel[None, None, :]).any(1).astype(np.float32) if sel_hit.any() else np.zeros((n, S), np.float32)
        real_sel = self._mat[users[:, None], self._sel[None, :]].astype(np.float32)
        tsel, msel = real_sel * fm_sel, fm_sel + np.float32(1)
        rsel = real_sel * msel
        nb = (n + batch - 1) // batch
        colptr = np.empty((nb, I + 1), dtype=np.int32)
        ent = np.empty(n * F, dtype=np.int32)
        L = _lib.lib()
        _lib.check(L.recad_aush_plan_columns(_vp(cols), n, batch, F, I, _vp(colptr), _vp(ent)), "recad_aush_plan_columns")
        dev = self.device
        with torch.cuda.device(dev):
            ints = torch.from_numpy(np.concatenate([cols.ravel(), colptr.ravel(), ent])).to(dev)
            flts = torch.from_numpy(np.concatenate([a.ravel() for a in (tval, dval, rsel, tsel, msel, zr)]).astype(np.float32)).to(dev)
            work = torch.empty(L.recad_aush_work_floats(I, n, batch, S), dtype=torch.float32, device=dev)
            loss = torch.empty(4, dtype=torch.float64, device=dev)
            ep = _lib.AushEpoch()
            ep.n_rows, ep.batch = n, batch
            ip, fp = ints.data_ptr(), flts.data_ptr()
            ep.cols, ep.colptr, ep.ent = ip, ip + 4 * n * F, ip + 4 * (n * F + nb * (I + 1))
            ep.tval, ep.dval = fp, fp + 4 * n * F
            ep.rsel, ep.tsel, ep.msel, ep.zr = (fp + 4 * (2 * n * F + k * n * S) for k in range(4))
            st = self._state(work)
            _lib.check(L.recad_aush_train_epoch(C.byref(st), C.byref(ep), self._step, loss.data_ptr(), ops._stream(dev)),
                       "recad_aush_train_epoch")
            out = loss.cpu()
        self._step += nb
        return tuple(float(x) for x in out)

    def _user_filter(self, train_mat, targets):
        return self._eligible(train_mat, targets)

    # ------------------------------------------------------------------ aush.py:182-230
    def generate_fake(self, **kwargs):
        self._require_instance("generate_fake")
        targets = list(kwargs["target_id_list"])
        available_idx = np.random.permutation(self._eligible(self._mat, targets))
        idx = available_idx[np.random.randint(0, len(available_idx), self.attack_num)].astype(np.int64)
        cols, tval, _ = self._draw_batch(idx, targets, with_zr=False)
        A, S = len(idx), len(self._sel)
        dev = self.device
        with torch.cuda.device(dev):
            gen = torch.zeros((A, S), dtype=torch.float32, device=dev)
            c, t = torch.from_numpy(cols).to(dev), torch.from_numpy(tval).to(dev)
            st = self._state()
            _lib.check(_lib.lib().recad_aush_generate(C.byref(st), c.data_ptr(), t.data_ptr(), A, gen.data_ptr(), ops._stream(dev)),
                       "recad_aush_generate")
            gen = gen.cpu().numpy()
        fake = np.zeros((A, self.n_items), dtype=np.float32)
        fake[np.arange(A)[:, None], cols] = self._mat[idx[:, None], cols]      # input_template = real * fillers_mask
        fake[:, self._sel] = fake[:, self._sel] + gen                           # + gen_output * selects_mask
        tg = np.unique(np.asarray(targets, dtype=np.int64))
        fake[:, tg] = fake[:, tg] + np.float32(5)                               # + target_patch
        patches = np.round(fake[:, self.selected_ids])
        patches[patches > 5] = 5
        patches[patches < 1] = 1
        fake[:, self.selected_ids] = patches
        return fake
