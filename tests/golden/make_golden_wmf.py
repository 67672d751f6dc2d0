"""Golden run of the reference's WMF surrogate trainer (recad/model/attacker/aia.py:393-489, unmodified) for the epochs
the reference CAN run here.  `fit_adv` imports `higher` unconditionally (aia.py:434) and the package is absent, so the
import is satisfied by a stub whose context manager hands back the model itself; with unroll_steps = 0 the stub is never
exercised arithmetically (the unrolled loop `range(epoch_num + 1, epoch_num + 1)` is empty) and every number stored
here is produced by reference code: the plain Adam epochs and the final P Q^T.

    cd <scratch dir>; PYTHONPATH=/root/reference python /root/repo/tests/golden/make_golden_wmf.py
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

stub = types.ModuleType("higher")


@contextlib.contextmanager
def innerloop_ctx(model, opt, *a, **k):
    yield model, None


stub.innerloop_ctx = innerloop_ctx
sys.modules["higher"] = stub

import recad  # noqa: E402,F401
from recad.model.attacker.aia import WMFTrainer, WeightedMF  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(11)
cases = {}
for name, (n_rows, n_items, dim, batch, epochs, wpos, wneg, wd) in {
        "a": (70, 40, 16, 16, 3, 1.0, 0.0, 1e-5),          # the defaults of default.py:177-184 (weight_neg 0: observed entries only)
        "b": (37, 53, 8, 5, 2, 2.0, 0.5, 1e-3)}.items():   # ragged last batch, both weights, stronger decay
    data = (rng.random((n_rows, n_items)) < 0.15) * rng.integers(1, 6, (n_rows, n_items))
    data = data.astype(np.float32)
    data[-5:] = rng.random((5, n_items)).astype(np.float32) * 5 * (rng.random((5, n_items)) < 0.6)     # "fake" rows: continuous
    torch.manual_seed(7)
    ref_init = WeightedMF(n_rows, n_items, dim)
    P0, Q0 = ref_init.P.detach().clone().numpy(), ref_init.Q.detach().clone().numpy()
    torch.manual_seed(7)
    np.random.seed(5)
    st = np.random.get_state()
    tr = WMFTrainer(n_users=n_rows, n_items=n_items, device=torch.device("cpu"), hidden_dim=dim, lr=1e-2, weight_decay=wd,
                    batch_size=batch, weight_pos=wpos, weight_neg=wneg)
    pred = tr.fit_adv(torch.tensor(data, requires_grad=True), epoch_num=epochs, unroll_steps=0)
    cases[name] = dict(data=data, P0=P0, Q0=Q0, pred=pred.detach().numpy(), P=tr.net.P.detach().numpy(), Q=tr.net.Q.detach().numpy(),
                       np_key=st[1], np_pos=st[2], hp=np.array([dim, batch, epochs, wpos, wneg, wd], dtype=np.float64))
np.savez_compressed(os.path.join(OUT, "wmf_plain.npz"), **{f"{c}_{k}": v for c, d in cases.items() for k, v in d.items()})
print({c: float(np.abs(d["pred"]).mean()) for c, d in cases.items()})
