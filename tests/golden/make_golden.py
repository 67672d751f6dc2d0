"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (the reference checkout does not exist on the
GPU box):

    mkdir -p /tmp/refrun/data && cp -r /root/reference/data/dev /tmp/refrun/data/
    python -c "import zipfile; zipfile.ZipFile('/root/reference/data/game.zip').extractall('/tmp/refrun/data')"
    cd /tmp/refrun && PYTHONPATH=/root/reference:/root/repo python /root/repo/tests/golden/make_golden.py

The reference writes caches relative to cwd (recad/utils.py:84-87), hence the
scratch cwd.  Everything stored here is an OUTPUT of unmodified reference code
(recad.dataset / recad.model.victim / recad.workflow) on seeded inputs; the
oracle (oracle/) and the CUDA path are both checked against these files.
"""
import hashlib
import io
import contextlib
import json
import os
import sys

import numpy as np
import torch

import recad  # the live reference  (seeds np.random / torch with 2023 on import)
from recad.dataset.implicit import pairwise_sample, pointwise_sample, shuffle
from recad.workflow import normal as ref_normal

sys.path.insert(0, "/root/repo")
from recad_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
recad.utils.TQDM = False
ref_normal.tqdm = recad.utils.tqdm


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def dict_flat(d):
    keys = np.array(list(d.keys()), dtype=np.int64)
    indptr = np.zeros(len(keys) + 1, dtype=np.int64)
    np.cumsum([len(d[k]) for k in d], out=indptr[1:])
    items = np.array([i for k in d for i in d[k]], dtype=np.int64)
    return keys, indptr, items


def graph_csr(ds):
    g = ds.Graph.to_sparse_csr()
    return (g.crow_indices().numpy().astype(np.int64), g.col_indices().numpy().astype(np.int64),
            g.values().numpy().astype(np.float32))


def allpos_flat(ds):
    indptr = np.zeros(ds.n_users + 1, dtype=np.int64)
    np.cumsum([len(p) for p in ds.allPos], out=indptr[1:])
    return indptr, np.concatenate([np.asarray(p, dtype=np.int64) for p in ds.allPos])


meta = {"versions": {"torch": torch.__version__, "numpy": np.__version__}}

# ------------------------------------------------------------------ datasets
dev = recad.dataset.from_config("implicit", "dev", need_graph=True, device=torch.device("cpu"))
dev_pt = recad.dataset.from_config("implicit", "dev", need_graph=False, sample="pointwise", device=torch.device("cpu"))
dev_tr = recad.dataset.from_config("implicit", "dev", need_graph=True, device=torch.device("cpu"),
                                   train_dict=dev.train_dict, valid_dict=dev.valid_dict, test_dict=dev.train_dict)
save = {}
for name, d in (("train", dev.train_dict), ("valid", dev.valid_dict), ("test", dev.test_dict)):
    k, p, i = dict_flat(d)
    save[f"{name}_keys"], save[f"{name}_indptr"], save[f"{name}_items"] = k, p, i
np.savez_compressed(os.path.join(OUT, "dev_dicts.npz"), **save)
meta["dev"] = {"n_users": dev.n_users, "n_items": dev.n_items, "train": dev.traindataSize,
               "valid": dev.validDataSize, "test": dev.testDataSize}

crow, col, val = graph_csr(dev)
crow_t, col_t, val_t = graph_csr(dev_tr)
ap_ptr, ap_idx = allpos_flat(dev)
ap_ptr_t, ap_idx_t = allpos_flat(dev_tr)
np.savez_compressed(os.path.join(OUT, "dev_graph.npz"), crow=crow, col=col, val=val,
                    crow_train=crow_t, col_train=col_t, val_train=val_t,
                    allpos_indptr=ap_ptr, allpos_indices=ap_idx,
                    allpos_indptr_train=ap_ptr_t, allpos_indices_train=ap_idx_t)
meta["dev_graph"] = {"sha": sha(crow, col, val), "sum": float(val.astype(np.float64).sum()),
                     "sha_train": sha(crow_t, col_t, val_t)}

game = recad.dataset.from_config("implicit", "game", need_graph=True, device=torch.device("cpu"))
game_tr = recad.dataset.from_config("implicit", "game", need_graph=True, device=torch.device("cpu"),
                                    train_dict=game.train_dict, valid_dict=game.valid_dict, test_dict=game.train_dict)
save = {}
for name, d in (("train", game.train_dict), ("valid", game.valid_dict), ("test", game.test_dict)):
    k, p, i = dict_flat(d)
    save[f"{name}_keys"], save[f"{name}_indptr"], save[f"{name}_items"] = k.astype(np.int32), p.astype(np.int32), i.astype(np.int32)
np.savez_compressed(os.path.join(OUT, "game_dicts.npz"), **save)
g1, g2 = graph_csr(game), graph_csr(game_tr)
meta["game"] = {"n_users": game.n_users, "n_items": game.n_items, "train": game.traindataSize,
                "graph_sha": sha(*g1), "graph_nnz": int(len(g1[1])), "graph_sum": float(g1[2].astype(np.float64).sum()),
                "graph_train_sha": sha(*g2), "graph_train_nnz": int(len(g2[1])),
                "graph_train_sum": float(g2[2].astype(np.float64).sum())}

# ml1m-shaped synthetic: intended (train) graph through the reference's own dok/lil path
tr, va, te = synthetic.make_splits(synthetic.ML1M, seed=0)
ml = recad.dataset.from_config("implicit", "ml1m", download=False, need_graph=True, device=torch.device("cpu"),
                               train_dict=tr, valid_dict=va, test_dict=tr)
g3 = graph_csr(ml)
meta["ml1m_shaped"] = {"n_users": ml.n_users, "n_items": ml.n_items, "train": ml.traindataSize,
                       "graph_train_sha": sha(*g3), "graph_train_nnz": int(len(g3[1])),
                       "graph_train_sum": float(g3[2].astype(np.float64).sum()),
                       "max_degree": int(np.diff(g3[0]).max())}

# ------------------------------------------------------------------ samplers
np.random.seed(2023)
S = pairwise_sample(dev)
st_after_pair = np.random.get_state()
perm = np.arange(len(S))
np.random.shuffle(perm)
P = pointwise_sample(dev_pt, 4)
st_end = np.random.get_state()
np.random.seed(2023)
Sg = pairwise_sample(game)
permg = np.arange(len(Sg)); np.random.shuffle(permg)
np.random.seed(2023)
Pg = pointwise_sample(game, 4)
np.savez_compressed(os.path.join(OUT, "samplers.npz"), dev_pairwise=S, dev_perm=perm, dev_pointwise=P,
                    dev_state_key_after_pairwise=st_after_pair[1], dev_state_pos_after_pairwise=st_after_pair[2],
                    dev_state_key_end=st_end[1], dev_state_pos_end=st_end[2],
                    game_pairwise_head=Sg[:64], game_pointwise_head=Pg[:64])
meta["samplers"] = {"seed": 2023, "game_pairwise_sha": sha(Sg), "game_pairwise_shape": list(Sg.shape),
                    "game_perm_sha": sha(permg), "game_pointwise_sha": sha(Pg), "game_pointwise_shape": list(Pg.shape)}


# ------------------------------------------------------------------ LightGCN on dev (train graph), D=64, 2 epochs
def record_batches(ds):
    """Wrap generate_batch so the exact batches the reference trained on are kept."""
    store = []
    orig = ds.generate_batch

    def wrapped(**kw):
        for b in orig(**kw):
            store.append({k: v.clone() for k, v in b.items()})
            yield b
    ds.generate_batch = wrapped
    return store


torch.manual_seed(2023)
np.random.seed(2023)
lgn = recad.model.from_config("victim", "lightgcn", latent_dim_rec=64, device=torch.device("cpu")).I(dataset=dev_tr)
init_u, init_i = lgn.embedding_user.weight.detach().clone().numpy(), lgn.embedding_item.weight.detach().clone().numpy()
batches = record_batches(dev_tr)
losses = [lgn.train_step()[0] for _ in range(2)]
lgn.eval()
qu = np.arange(0, 512, 7, dtype=np.int64)
qi = (qu * 3 + 1) % dev_tr.n_items
with torch.no_grad():
    fwd = lgn(torch.as_tensor(qu), torch.as_tensor(qi)).numpy()
    ou, oi = lgn.computer()
np.savez_compressed(os.path.join(OUT, "lightgcn_dev.npz"), init_user=init_u, init_item=init_i,
                    final_user=lgn.embedding_user.weight.detach().numpy(), final_item=lgn.embedding_item.weight.detach().numpy(),
                    out_user=ou.numpy(), out_item=oi.numpy(),
                    losses=np.array(losses, dtype=np.float64), q_users=qu, q_items=qi, q_scores=fwd,
                    batch_users=torch.cat([b["users"] for b in batches]).numpy(),
                    batch_pos=torch.cat([b["positive_items"] for b in batches]).numpy(),
                    batch_neg=torch.cat([b["negative_items"] for b in batches]).numpy(),
                    batch_sizes=np.array([len(b["users"]) for b in batches], dtype=np.int64))
meta["lightgcn_dev"] = {"D": 64, "L": 3, "lambda": 1e-4, "lr": 1e-3, "epochs": 2, "losses": losses,
                        "graph": "dev train edges (test_dict=train_dict)", "batch": 1024}

# ------------------------------------------------------------------ MF on dev, d=64, 2 epochs
torch.manual_seed(2023)
np.random.seed(2023)
mf = recad.model.from_config("victim", "mf", embedding_size=64, device=torch.device("cpu")).I(dataset=dev_pt)
mf_init = [p.detach().clone().numpy() for p in (mf.user_emb.weight, mf.user_bias.weight, mf.item_emb.weight, mf.item_bias.weight)]
batches = record_batches(dev_pt)
mf_losses = [mf.train_step()[0] for _ in range(2)]
mf_final = [p.detach().clone().numpy() for p in (mf.user_emb.weight, mf.user_bias.weight, mf.item_emb.weight, mf.item_bias.weight)]
with torch.no_grad():
    mf_fwd = mf(torch.as_tensor(qu), torch.as_tensor(qi)).numpy()
np.savez_compressed(os.path.join(OUT, "mf_dev.npz"),
                    **{f"init{k}": a for k, a in enumerate(mf_init)}, **{f"final{k}": a for k, a in enumerate(mf_final)},
                    losses=np.array(mf_losses), q_users=qu, q_items=qi, q_scores=mf_fwd,
                    batch_users=torch.cat([b["users"] for b in batches]).numpy(),
                    batch_items=torch.cat([b["items"] for b in batches]).numpy(),
                    batch_labels=torch.cat([b["labels"] for b in batches]).numpy(),
                    batch_sizes=np.array([len(b["users"]) for b in batches], dtype=np.int64))
del dev_pt.generate_batch
meta["mf_dev"] = {"embedding_size": 64, "mean": 3.0, "lr": 1e-3, "epochs": 2, "losses": mf_losses, "batch": 1024}

# ------------------------------------------------------------------ NCF on dev, small tower (f=8, L=3), 2 epochs
torch.manual_seed(2023)
np.random.seed(2023)
ncf = recad.model.from_config("victim", "ncf", factor_num=8, num_layers=3, device=torch.device("cpu")).I(dataset=dev_pt)


def ncf_params(m):
    lin = [l for l in m.MLP_layers if isinstance(l, torch.nn.Linear)]
    d = {"ug": m.embed_user_GMF.weight, "ig": m.embed_item_GMF.weight, "um": m.embed_user_MLP.weight,
         "im": m.embed_item_MLP.weight, "Wp": m.predict_layer.weight, "bp": m.predict_layer.bias}
    for k, l in enumerate(lin):
        d[f"W{k}"], d[f"b{k}"] = l.weight, l.bias
    return {k: v.detach().clone().numpy() for k, v in d.items()}


ncf_init = ncf_params(ncf)
batches = record_batches(dev_pt)
ncf_losses = [ncf.train_step()[0] for _ in range(2)]
ncf_final = ncf_params(ncf)
with torch.no_grad():
    ncf_fwd = ncf(torch.as_tensor(qu), torch.as_tensor(qi)).numpy()
np.savez_compressed(os.path.join(OUT, "ncf_dev.npz"),
                    **{f"init_{k}": a for k, a in ncf_init.items()}, **{f"final_{k}": a for k, a in ncf_final.items()},
                    losses=np.array(ncf_losses), q_users=qu, q_items=qi, q_scores=ncf_fwd,
                    batch_users=torch.cat([b["users"] for b in batches]).numpy(),
                    batch_items=torch.cat([b["items"] for b in batches]).numpy(),
                    batch_labels=torch.cat([b["labels"] for b in batches]).numpy(),
                    batch_sizes=np.array([len(b["users"]) for b in batches], dtype=np.int64))
del dev_pt.generate_batch
meta["ncf_dev"] = {"factor_num": 8, "num_layers": 3, "lr": 1e-3, "epochs": 2, "losses": ncf_losses, "batch": 1024}

# default-size NCF (f=32, L=5): too large to store; losses + a few rows only.  Initial
# weights are re-creatable on the same torch build: torch.manual_seed(2023) followed by the
# module constructor (ncf.py:32-77).
torch.manual_seed(2023)
np.random.seed(2023)
ncf_big = recad.model.from_config("victim", "ncf", device=torch.device("cpu")).I(dataset=dev_pt)
big_init_probe = ncf_params(ncf_big)
big_losses = [ncf_big.train_step()[0] for _ in range(1)]
big_final = ncf_params(ncf_big)
with torch.no_grad():
    big_fwd = ncf_big(torch.as_tensor(qu), torch.as_tensor(qi)).numpy()
np.savez_compressed(os.path.join(OUT, "ncf_dev_default.npz"), losses=np.array(big_losses), q_users=qu, q_items=qi,
                    q_scores=big_fwd, init_um_rows=big_init_probe["um"][:4], final_um_rows=big_final["um"][:4],
                    final_W4=big_final["W4"], final_Wp=big_final["Wp"], init_W4=big_init_probe["W4"])
meta["ncf_dev_default"] = {"factor_num": 32, "num_layers": 5, "epochs": 1, "losses": big_losses}


# ------------------------------------------------------------------ evaluation (normal.py:57-160)
class _Cap:
    rows = None


def run_eval(model_a, model_b, ds, targets, topks):
    wf = object.__new__(ref_normal.Normal)
    wf.c = {"device": torch.device("cpu")}
    wf.logger = recad.utils.get_logger("golden")
    captured = {}
    orig = ref_normal.fmt_tab
    ref_normal.fmt_tab = lambda table, **kw: captured.setdefault("table", table) and ""
    gen = wf.user_item_model_generate
    rows = []

    def cap_gen(*a, **k):
        r = gen(*a, **k)
        rows.append(r.copy())
        return r
    wf.user_item_model_generate = cap_gen
    with contextlib.redirect_stdout(io.StringIO()):
        wf.normal_evaluate(model_a, model_b, ds, targets, topks)
    ref_normal.fmt_tab = orig
    return dict(captured["table"]), rows


topks = [10, 20, 50, 100]
# MF: clean = a freshly initialised MF, "fake" = the trained one above
torch.manual_seed(7)
mf_a = recad.model.from_config("victim", "mf", embedding_size=64, device=torch.device("cpu")).I(dataset=dev_pt)
mf_a_w = [p.detach().clone().numpy() for p in (mf_a.user_emb.weight, mf_a.user_bias.weight, mf_a.item_emb.weight, mf_a.item_bias.weight)]
tab, rows = run_eval(mf_a, mf, dev_pt, [0], topks)
tab5, rows5 = run_eval(mf_a, mf, dev_pt, [5], topks)
np.savez_compressed(os.path.join(OUT, "eval_mf_dev.npz"), **{f"a{k}": a for k, a in enumerate(mf_a_w)},
                    rows_a=rows[0], rows_b=rows[1], rows5_a=rows5[0], rows5_b=rows5[1])
meta["eval_mf_dev"] = {"targets": [0], "topks": topks, "table": tab, "table_target5": tab5}

# LightGCN (reference-actual graph, i.e. the test-split bug), clean = init, fake = trained 1 epoch
torch.manual_seed(11)
np.random.seed(2023)
lg_a = recad.model.from_config("victim", "lightgcn", latent_dim_rec=64, device=torch.device("cpu")).I(dataset=dev)
lg_b = recad.model.from_config("victim", "lightgcn", latent_dim_rec=64, device=torch.device("cpu")).I(dataset=dev)
lg_b.train_step()
tabl, rowsl = run_eval(lg_a, lg_b, dev, [3], topks)
np.savez_compressed(os.path.join(OUT, "eval_lightgcn_dev.npz"),
                    a_user=lg_a.embedding_user.weight.detach().numpy(), a_item=lg_a.embedding_item.weight.detach().numpy(),
                    b_user=lg_b.embedding_user.weight.detach().numpy(), b_item=lg_b.embedding_item.weight.detach().numpy(),
                    rows_a=rowsl[0], rows_b=rowsl[1])
meta["eval_lightgcn_dev"] = {"targets": [3], "topks": topks, "table": tabl, "graph": "reference-actual (test split)"}

# ------------------------------------------------------------------ whole workflow (normal.py:162-225)
def run_workflow(victim, **vkw):
    import random
    random.seed(2023); np.random.seed(2023); torch.manual_seed(2023)
    sampling = "pointwise" if victim in ("mf", "ncf") else "pairwise"
    cfg = {
        "victim_data": recad.dataset.from_config("implicit", "dev", need_graph=victim == "lightgcn", sample=sampling,
                                                 device=torch.device("cpu")),
        "attack_data": recad.dataset.from_config("explicit", "dev", device=torch.device("cpu")).partial_sample(user_ratio=0.2),
        "victim": recad.model.from_config("victim", victim, device=torch.device("cpu"), **vkw),
        "attacker": recad.model.from_config("attacker", "random", filler_num=36, device=torch.device("cpu")),
        "rec_epoch": 2, "attack_epoch": 1, "device": torch.device("cpu"),
    }
    wf = recad.workflow.from_config("no defense", **cfg)
    rec = {}
    gen_fake = wf.attacker.generate_fake

    def cap_fake(**kw):
        fa = gen_fake(**kw)
        rec["fake"] = fa.copy()
        rec["np_state_after_fake"] = np.random.get_state()
        rec["torch_state_after_fake"] = torch.get_rng_state().clone()
        return fa
    wf.attacker.generate_fake = cap_fake
    rec["np_state_start"] = np.random.get_state()
    rec["torch_state_start"] = torch.get_rng_state().clone()
    rec["init"] = {k: v.detach().clone().numpy() for k, v in wf.victim.state_dict().items()}
    captured = {}
    orig = ref_normal.fmt_tab
    ref_normal.fmt_tab = lambda table, **kw: captured.setdefault("table", table) and ""
    with contextlib.redirect_stdout(io.StringIO()):
        wf.execute()
    ref_normal.fmt_tab = orig
    rec["table"] = dict(captured["table"])
    rec["final"] = {k: v.detach().clone().numpy() for k, v in wf.victim.state_dict().items()}
    return rec


for victim, kw in (("mf", {"embedding_size": 64}), ("lightgcn", {"latent_dim_rec": 64})):
    rec = run_workflow(victim, **kw)
    fr, fc = np.nonzero(rec["fake"])
    np.savez_compressed(os.path.join(OUT, f"workflow_{victim}_dev.npz"),
                        fake_shape=np.array(rec["fake"].shape), fake_rows=fr, fake_cols=fc, fake_vals=rec["fake"][fr, fc],
                        np_key_start=rec["np_state_start"][1], np_pos_start=rec["np_state_start"][2],
                        np_key_after_fake=rec["np_state_after_fake"][1], np_pos_after_fake=rec["np_state_after_fake"][2],
                        torch_state_start=rec["torch_state_start"].numpy(),
                        torch_state_after_fake=rec["torch_state_after_fake"].numpy(),
                        **{f"init__{k}": v for k, v in rec["init"].items()},
                        **{f"final__{k}": v for k, v in rec["final"].items()})
    meta[f"workflow_{victim}_dev"] = {"rec_epoch": 2, "attack": "random", "table": rec["table"], "victim_kwargs": kw}

with open(os.path.join(OUT, "meta.json"), "w") as f:
    json.dump(meta, f, indent=1, sort_keys=True)
print(json.dumps(meta, indent=1)[:3000])
