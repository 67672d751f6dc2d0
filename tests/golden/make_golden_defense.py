"""Golden fixture of the "defense" workflow (recad/workflow/defense.py:175-303) from the LIVE reference.

    mkdir -p /tmp/refrun/data && cp -r /root/reference/data/dev /tmp/refrun/data/
    cd /tmp/refrun && PYTHONPATH=/root/reference:/root/repo python /root/repo/tests/golden/make_golden_defense.py

The reference's only defender (PCASelectUsers.py:50) allocates on 'cuda' unconditionally and cannot run in the
CPU build container, and the defender is outside the hot path anyway: a duck-typed fixture defender with a FIXED
answer (30 of the 50 fake users + 5 genuine users) stands in for it, so what is pinned is the workflow's own
sequence -- attack, re-seed, retrain, evaluate, delete_data, re-seed, retrain, evaluate -- with unmodified
reference datasets, MF victim, random attacker and evaluator.  Separate from make_golden.py so that the other
fixtures (which depend on the global RNG order of that script) stay byte-identical.
"""
import contextlib
import io
import json
import os
import random

import numpy as np
import torch

import recad
from recad.workflow import defense as ref_defense

OUT = os.path.dirname(os.path.abspath(__file__))
recad.utils.TQDM = False
ref_defense.tqdm = recad.utils.tqdm
CPU = torch.device("cpu")


class FixtureDefender:
    model_name = "fixture_defender"

    def __init__(self, flagged):
        self.flagged = list(flagged)

    def I(self, **kw):       # noqa: E743 -- the reference's instantiation hook
        return self

    def to(self, device):
        return self

    def input_describe(self):
        return {"defense_step": {}}

    def defense_step(self, **kw):
        return list(self.flagged)


random.seed(2023); np.random.seed(2023); torch.manual_seed(2023)
victim_data = recad.dataset.from_config("implicit", "dev", need_graph=False, sample="pointwise", device=CPU)
U = victim_data.n_users
flagged = [U + r for r in range(0, 50) if r % 5 != 0 and r < 38][:30] + [3, 17, 40, 99, 123]
cfg = {
    "victim_data": victim_data,
    "attack_data": recad.dataset.from_config("explicit", "dev", device=CPU).partial_sample(user_ratio=0.2),
    "defense_data": victim_data,
    "victim": recad.model.from_config("victim", "mf", device=CPU, embedding_size=64),
    "attacker": recad.model.from_config("attacker", "random", filler_num=36, device=CPU),
    "defender": FixtureDefender(flagged),
    "rec_epoch": 2, "attack_epoch": 1, "device": CPU,
}
wf = recad.workflow.from_config("defense", **cfg)
rec = {}
gen_fake = wf.attacker.generate_fake


def cap_fake(**kw):
    fa = gen_fake(**kw)
    rec["fake"] = fa.copy()
    return fa


wf.attacker.generate_fake = cap_fake
delete = victim_data.delete_data


def cap_delete(*a, **k):
    ds = delete(*a, **k)
    rec["cleaned_n_users"] = int(ds.n_users)
    rec["cleaned_train_size"] = int(ds.traindataSize)
    return ds


victim_data.delete_data = cap_delete
rec["np_state_start"] = np.random.get_state()
rec["init"] = {k: v.detach().clone().numpy() for k, v in wf.victim.state_dict().items()}
tables = []
orig = ref_defense.fmt_tab
ref_defense.fmt_tab = lambda table, **kw: tables.append(dict(table)) or ""
with contextlib.redirect_stdout(io.StringIO()):
    wf.execute()
ref_defense.fmt_tab = orig
assert len(tables) == 2, len(tables)
fr, fc = np.nonzero(rec["fake"])
np.savez_compressed(os.path.join(OUT, "workflow_defense_mf_dev.npz"),
                    fake_shape=np.array(rec["fake"].shape), fake_rows=fr, fake_cols=fc, fake_vals=rec["fake"][fr, fc],
                    np_key_start=rec["np_state_start"][1], np_pos_start=rec["np_state_start"][2],
                    flagged=np.array(flagged, dtype=np.int64),
                    **{f"init__{k}": v for k, v in rec["init"].items()})
meta = {"rec_epoch": 2, "attack": "random", "victim_kwargs": {"embedding_size": 64}, "n_users": int(U),
        "table_after_attack": tables[0], "table_after_defense": tables[1],
        "cleaned_n_users": rec["cleaned_n_users"], "cleaned_train_size": rec["cleaned_train_size"]}
with open(os.path.join(OUT, "meta_defense.json"), "w") as f:
    json.dump(meta, f, indent=1, sort_keys=True)
print(json.dumps(meta, indent=1))
