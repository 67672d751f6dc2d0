"""BASELINE.json configs[1] as stated: LightGCN (3 layers, dim 64) on the ml1m shape with AUSH fake-profile injection,
rec_epoch = 20, one B200 -- run through the REFERENCE's own workflow: the unmodified `recad` package (baseline/_ref)
builds `Normal` with its own AUSH attacker and explicit attack dataset; `recad_b200.register.install(override=True)` has
rebound the victim, the implicit dataset and the evaluator to the CUDA path.  (ml1m.zip is absent: shape-matched
synthetic interactions with ratings 1-5; the implicit side keeps ratings >= 4 like implicit.py:94-127.)

    python tools/config2_execute.py [--rec-epoch 20] [--attack-epoch 5]

Prints one JSON line: wall-clock of execute() and of its phases, the evaluation table.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rec-epoch", type=int, default=20)
    ap.add_argument("--attack-epoch", type=int, default=5)
    ap.add_argument("--attacker", choices=["b200", "reference"], default="b200",
                    help="b200: the registered CUDA Aush + explicit dataset; reference: the reference's own classes")
    a = ap.parse_args()
    os.chdir(tempfile.mkdtemp())                       # the reference resolves ./data and ./generated against cwd
    sys.path.insert(0, REF)
    import recad
    import recad_b200.register as reg
    from recad_b200 import synthetic
    reg.install(override=True)
    recad.utils.TQDM = False
    dev = torch.device("cuda:0")
    import random
    random.seed(2023); np.random.seed(2023); torch.manual_seed(2023)
    tr, va, te = synthetic.make_splits(synthetic.ML1M, seed=0)
    U, I = synthetic.ML1M["n_users"], synthetic.ML1M["n_items"]
    rng = np.random.default_rng(1)

    def kvr(d, lo):                                    # explicit rows (user, item, rating); implicit positives are the ratings >= 4
        rows = [(u, i, int(rng.integers(lo, 6))) for u, items in d.items() for i in items]
        return np.asarray(rows, dtype=np.int64)
    ex_train = np.concatenate([kvr(tr, 4), kvr({u: rng.choice(I, 20).tolist() for u in range(0, U, 3)}, 1)])   # + some low ratings
    t0 = time.time()
    victim_data = recad.dataset.from_config("implicit", "ml1m", need_graph=True, sample="pairwise", device=dev, download=False,
                                            train_dict=tr, valid_dict=va, test_dict=te, graph_edges="train")
    explicit_cls = recad.dataset.factories["explicit"] if a.attacker == "b200" else recad.dataset.explicit.ExplicitData
    aush_cls = recad.model.factories["attacker"]["aush"] if a.attacker == "b200" else recad.model.attacker.Aush
    attack_data = explicit_cls.from_config("ml1m", device=dev, download=False, train_dict=ex_train.astype(np.float64),
                                           valid_dict=kvr(va, 4).astype(np.float64), test_dict=kvr(te, 4).astype(np.float64)
                                           ).partial_sample(user_ratio=0.2)
    cfg = {"victim_data": victim_data, "attack_data": attack_data,
           "victim": recad.model.from_config("victim", "lightgcn", latent_dim_rec=64, lightGCN_n_layers=3, device=dev),
           "attacker": aush_cls.from_config(device=dev),
           "rec_epoch": a.rec_epoch, "attack_epoch": a.attack_epoch, "device": dev}
    wf = recad.workflow.from_config("no defense", **cfg)
    t_build = time.time() - t0
    assert type(wf).__module__ == "recad.workflow.normal"
    assert type(wf.attacker).__module__.startswith("recad." if a.attacker == "reference" else "recad_b200.")
    phases, seen = {}, {}
    nt = type(wf).normal_train

    def timed_train(self, **kw):
        torch.cuda.synchronize(); t = time.time()
        out = nt(self, **kw)
        torch.cuda.synchronize()
        key = f"attacker_train_s ({a.attacker} AUSH)" if kw["model"] is self.attacker else f"victim_train_s[{len([k for k in phases if k.startswith('victim')])}]"
        phases[key] = round(time.time() - t, 3)
        return out
    ev = type(wf).normal_evaluate

    def timed_eval(self, *args, **kw):
        torch.cuda.synchronize(); t = time.time()
        seen["table"] = ev(self, *args, **kw)
        torch.cuda.synchronize()
        phases["evaluate_s"] = round(time.time() - t, 4)
        return seen["table"]
    type(wf).normal_train, type(wf).normal_evaluate = timed_train, timed_eval
    torch.cuda.synchronize(); t0 = time.time()
    wf.execute()                                       # recad/workflow/normal.py:162-225, unmodified
    torch.cuda.synchronize()
    total = time.time() - t0
    print(json.dumps({"tool": "config2_execute", "workload": f"ml1m-shaped {U} x {I}, {synthetic.ML1M['train']} train interactions; LightGCN D=64 L=3, "
                      f"rec_epoch={a.rec_epoch}, AUSH attacker ({a.attacker} classes, attack_epoch={a.attack_epoch}), 50 fake users",
                      "execute_s": round(total, 3), "build_s (datasets + graph + models)": round(t_build, 3), "phases": phases,
                      "victim_epoch_s": round(phases.get("victim_train_s[0]", 0) / max(a.rec_epoch, 1), 4),
                      "table": {k: float(v) for k, v in seen["table"].items()}}))


if __name__ == "__main__":
    main()
