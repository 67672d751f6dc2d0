"""Epoch / evaluation timings of the three victims on the ml1m-shaped graph (BASELINE.json configs[0..2] shapes), or of
NCF on the yelp-shaped one (configs[2]: 54 632 x 34 474, 1.5 M interactions): one JSON line per model, with the CPU
oracle port (torch CPU, all host threads) timed on a bounded sample beside it.
    python tools/victims_bench.py > profiles/victims_ml1m_r02.jsonl
    python tools/victims_bench.py yelp > profiles/ncf_yelp_r02.jsonl"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pointwise_models as opm  # noqa: E402  (CPU baseline leg only)
from recad_b200 import dataset, evaluate, model, synthetic  # noqa: E402

DEV = torch.device("cuda:0")
YELP = len(sys.argv) > 1 and sys.argv[1] == "yelp"
SHAPE = synthetic.YELP if YELP else synthetic.ML1M
tr, va, te = synthetic.make_splits(SHAPE, seed=0)
torch.set_num_threads(os.cpu_count() or 1)


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / n, out


def cpu_epoch_estimate(oracle, samples, batch, n_batches_total, n_time=20):
    t0 = time.time()
    for k in range(n_time):
        s = samples[k * batch:(k + 1) * batch]
        oracle.step(s[:, 0], s[:, 1], s[:, 2])
    return (time.time() - t0) / n_time * n_batches_total


MODELS = ((("ncf", {}), ("ncf", {"tower_precision": "fp32"})) if YELP else
          (("mf", {"embedding_size": 64}), ("ncf", {}), ("ncf", {"tower_precision": "fp32"}), ("lightgcn", {"latent_dim_rec": 64})))
for name, kw in MODELS:
    pairwise = name == "lightgcn"
    data = dataset.from_config("implicit", "yelp" if YELP else "ml1m", train_dict=tr, valid_dict=va, test_dict=te, need_graph=pairwise,
                               graph_edges="train", sample="pairwise" if pairwise else "pointwise", device=DEV)
    torch.manual_seed(2023)
    np.random.seed(2023)
    m = model.from_config("victim", name, device=DEV, **kw).I(dataset=data)
    ep_s, loss = timed(lambda: m.train_step())
    if YELP:      # NCF full ranking = one tower evaluation per (user, item) pair (2.6 PFLOP for all users): time 4096 users
        some = evaluate.eligible_users(data, [0])[:4096]
        ev_s, rows = timed(lambda: evaluate.model_rows(m, data, [0], [10, 20, 50, 100], users=some)[0], n=1)
    else:
        ev_s, rows = timed(lambda: evaluate.model_rows(m, data, [0], [10, 20, 50, 100])[0])
    n = data.traindataSize * (1 if pairwise else 5)
    B = 1024
    rec = {"victim": name, **{k: v for k, v in kw.items()},
           "workload": f"{'yelp' if YELP else 'ml1m'}-shaped {SHAPE['n_users']} x {SHAPE['n_items']}, {SHAPE['train']} train interactions",
           "samples_per_epoch": n, "batch": B, "epoch_s (train_step, incl. exact host sampler + H2D)": round(ep_s, 4),
           "loss": loss[0], "fullrank_eval_s (all eligible users, HR@k rows)": round(ev_s, 5), "eval_users": int(len(rows))}
    if not pairwise:
        samples, perm = data.epoch_samples(DEV)
        S = samples[perm].cpu().numpy()
        torch.manual_seed(2023)
        if name == "mf":
            o = opm.MFOracle(*(p.detach().cpu().numpy() for p in (m.user_emb.weight, m.user_bias.weight, m.item_emb.weight, m.item_bias.weight)))
        else:
            lin = [l for l in m.MLP_layers if isinstance(l, torch.nn.Linear)]
            o = opm.NCFOracle({"ug": m.embed_user_GMF.weight.cpu().numpy(), "ig": m.embed_item_GMF.weight.cpu().numpy(),
                               "um": m.embed_user_MLP.weight.cpu().numpy(), "im": m.embed_item_MLP.weight.cpu().numpy(),
                               "W": [l.weight.cpu().numpy() for l in lin], "b": [l.bias.cpu().numpy() for l in lin],
                               "Wp": m.predict_layer.weight.cpu().numpy(), "bp": m.predict_layer.bias.cpu().numpy()})
        rec["cpu_port_epoch_s (extrapolated from 20 batches, %d threads)" % torch.get_num_threads()] = round(
            cpu_epoch_estimate(o, S, B, (n + B - 1) // B), 2)
    print(json.dumps(rec), flush=True)
