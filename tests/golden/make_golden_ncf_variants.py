"""Golden runs of the reference NCF's other variants (recad/model/victim/ncf.py:49-53, 62-104, 112-131): 'GMF', 'MLP' and
'NeuMF-pre' (initialised from the two trained ones), on the recorded batches of ncf_dev.npz through a stub dataset.

    cd <scratch dir with data/dev>; PYTHONPATH=/root/reference python /root/repo/tests/golden/make_golden_ncf_variants.py
"""
import os
import sys

import numpy as np
import torch

import recad  # noqa: F401

sys.path.insert(0, "/root/repo")
from tests import util  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
z = util.load("ncf_dev.npz")
meta = util.meta()
U, I = meta["dev"]["n_users"], meta["dev"]["n_items"]
batches = util.split_batches(z, ("batch_users", "batch_items", "batch_labels"))
per_epoch = len(batches) // 2


class Stub:
    dataset_name = "stub"

    def __init__(self):
        self.epoch = 0

    def info_describe(self):
        return {"n_users": U, "n_items": I}

    def generate_batch(self, **kw):
        lo = (self.epoch % 2) * per_epoch
        self.epoch += 1
        for u, i, y in batches[lo:lo + per_epoch]:
            yield {"users": torch.as_tensor(u), "items": torch.as_tensor(i), "labels": torch.as_tensor(y)}


def params(m):
    lin = [l for l in m.MLP_layers if isinstance(l, torch.nn.Linear)]
    d = {"ug": m.embed_user_GMF.weight, "ig": m.embed_item_GMF.weight, "um": m.embed_user_MLP.weight,
         "im": m.embed_item_MLP.weight, "Wp": m.predict_layer.weight, "bp": m.predict_layer.bias}
    for k, l in enumerate(lin):
        d[f"W{k}"], d[f"b{k}"] = l.weight, l.bias
    return {k: v.detach().clone().numpy() for k, v in d.items()}


out, models = {}, {}
qu, qi = z["q_users"], z["q_items"]
for name in ("GMF", "MLP"):
    torch.manual_seed(31)
    m = recad.model.from_config("victim", "ncf", factor_num=8, num_layers=3, model=name, device=torch.device("cpu")).I(dataset=Stub())
    out.update({f"{name}_init_{k}": v for k, v in params(m).items()})
    losses = [m.train_step()[0] for _ in range(2)]
    out.update({f"{name}_final_{k}": v for k, v in params(m).items()})
    out[f"{name}_losses"] = np.array(losses)
    with torch.no_grad():
        out[f"{name}_q_scores"] = m(torch.as_tensor(qu), torch.as_tensor(qi)).numpy()
    models[name] = m
torch.manual_seed(32)
pre = recad.model.from_config("victim", "ncf", factor_num=8, num_layers=3, model="NeuMF-pre", GMF_model=models["GMF"],
                              MLP_model=models["MLP"], device=torch.device("cpu")).I(dataset=Stub())
out.update({f"pre_init_{k}": v for k, v in params(pre).items()})
out["pre_losses"] = np.array([pre.train_step()[0]])
out.update({f"pre_final_{k}": v for k, v in params(pre).items()})
with torch.no_grad():
    out["pre_q_scores"] = pre(torch.as_tensor(qu), torch.as_tensor(qi)).numpy()
np.savez_compressed(os.path.join(OUT, "ncf_variants_dev.npz"), **out)
print({k: out[k].tolist() for k in out if k.endswith("losses")})
