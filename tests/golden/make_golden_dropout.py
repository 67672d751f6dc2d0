"""Golden fixture for LightGCN graph dropout (recad/model/victim/lightgcn.py:62-80, 90-96) from the LIVE reference.

    mkdir -p /tmp/refrun/data && cp -r /root/reference/data/dev /tmp/refrun/data/
    cd /tmp/refrun && PYTHONPATH=/root/reference:/root/repo python /root/repo/tests/golden/make_golden_dropout.py

The unmodified reference LightGCN (dropout = 1, keep_prob = 0.6, D = 32, 2 layers) on the dev train graph, CPU: two epochs
(batch 128) on recorded batches, then one forward in training mode (the reference never calls eval(): every forward of a workflow
draws a fresh mask) and one in eval mode.  The masks come from torch's CPU generator (`torch.rand(nnz)` per computer()
call), seeded with torch.manual_seed(77) BEFORE the model is built; the fixture keeps everything a replay needs.
"""
import os
import sys

import numpy as np
import torch

import recad  # the live reference

sys.path.insert(0, "/root/repo")
OUT = os.path.dirname(os.path.abspath(__file__))
recad.utils.TQDM = False

dev = recad.dataset.from_config("implicit", "dev", need_graph=True, device=torch.device("cpu"))
dev_tr = recad.dataset.from_config("implicit", "dev", need_graph=True, device=torch.device("cpu"), pairwise_batch_size=128,
                                   train_dict=dev.train_dict, valid_dict=dev.valid_dict, test_dict=dev.train_dict)


def record_batches(ds):
    store = []
    orig = ds.generate_batch

    def wrapped(**kw):
        for b in orig(**kw):
            store.append({k: v.clone() for k, v in b.items()})
            yield b
    ds.generate_batch = wrapped
    return store


torch.manual_seed(77)
np.random.seed(77)
lgn = recad.model.from_config("victim", "lightgcn", latent_dim_rec=32, lightGCN_n_layers=2, dropout=1, keep_prob=0.6,
                              device=torch.device("cpu")).I(dataset=dev_tr)
init_u, init_i = lgn.embedding_user.weight.detach().clone().numpy(), lgn.embedding_item.weight.detach().clone().numpy()
batches = record_batches(dev_tr)
losses = [lgn.train_step()[0] for _ in range(2)]
qu = np.arange(0, 512, 7, dtype=np.int64)
qi = (qu * 3 + 1) % dev_tr.n_items
with torch.no_grad():
    fwd_train = lgn(torch.as_tensor(qu), torch.as_tensor(qi)).numpy()          # training mode: a fresh mask
    lgn.eval()
    fwd_eval = lgn(torch.as_tensor(qu), torch.as_tensor(qi)).numpy()           # eval mode: the full graph
np.savez_compressed(os.path.join(OUT, "lightgcn_dropout_dev.npz"), init_user=init_u, init_item=init_i,
                    final_user=lgn.embedding_user.weight.detach().numpy(), final_item=lgn.embedding_item.weight.detach().numpy(),
                    losses=np.array(losses, dtype=np.float64), q_users=qu, q_items=qi, q_scores_train=fwd_train, q_scores_eval=fwd_eval,
                    batch_users=torch.cat([b["users"] for b in batches]).numpy(),
                    batch_pos=torch.cat([b["positive_items"] for b in batches]).numpy(),
                    batch_neg=torch.cat([b["negative_items"] for b in batches]).numpy(),
                    batch_sizes=np.array([len(b["users"]) for b in batches], dtype=np.int64),
                    nnz=np.int64(lgn.Graph._nnz()), seed=np.int64(77), keep_prob=np.float64(0.6))
print("losses", losses, "batches", len(batches), "nnz", lgn.Graph._nnz())
