"""Golden run of the reference's AUSH attacker (recad/model/attacker/aush.py, unmodified, CPU) on a synthetic explicit
dataset (the shipped `dev` set has at most 6 ratings per user, so no user passes the filler filter): 400 users x 300 items,
20-80 ratings of 1..5 per user, handed to the reference's own ExplicitData through its in-memory entry point
(explicit.py:47-52).  Stored: the rating rows, the generator / discriminator state before and after every `train_step`, the four epoch losses, the global
numpy / torch generator states at the start, and the fake profiles `generate_fake` draws afterwards.

Two cases: the default (one selected item, ZR_ratio 0.2, batch 256) and a harder one (three selected items, two targets,
batch 100 -> ragged last batch, ZR_ratio 0.5, filler_num 12).

    cd <scratch dir>
    PYTHONPATH=/root/reference python /root/repo/tests/golden/make_golden_aush.py
"""
import os
import random

import numpy as np
import torch

import recad

OUT = os.path.dirname(os.path.abspath(__file__))
cpu = torch.device("cpu")
cases = {
    "a": dict(ds={}, att={}, targets=[0], epochs=3),
    "b": dict(ds={"batch_size": 100}, att={"selected_ids": [62, 5, 140], "ZR_ratio": 0.5, "filler_num": 12, "attack_num": 20},
              targets=[7, 3], epochs=2),
}
store = {}
rng = np.random.default_rng(3)
rows = []
for u in range(400):
    items = rng.choice(300, size=int(rng.integers(20, 81)), replace=False)
    rows += [(u, int(i), int(r)) for i, r in zip(items, rng.integers(1, 6, len(items)))]
rows = np.asarray(rows, dtype=np.float64)            # load_file_as_np gives a float/int [n, 3] array; float is what remap keeps
perm = rng.permutation(len(rows))
n_te = len(rows) // 10
train, test = rows[perm[n_te:]], rows[perm[:n_te]]
train = train[np.lexsort((train[:, 1], train[:, 0]))]
store["train"], store["test"] = train.astype(np.int16), test.astype(np.int16)
for name, c in cases.items():
    random.seed(2023); np.random.seed(2023); torch.manual_seed(2023)
    ds = recad.dataset.from_config("explicit", "dev", device=cpu, if_cache=False, remap_enable=False, train_dict=train.copy(),
                                   valid_dict=test.copy(), test_dict=test.copy(), **c["ds"])
    att = recad.model.from_config("attacker", "aush", device=cpu, **c["att"]).I(dataset=ds)
    d = {}
    for k, v in att.netG.state_dict().items():
        d[f"G0__{k}"] = v.numpy().copy()
    for k, v in att.netD.state_dict().items():
        d[f"D0__{k}"] = v.numpy().copy()
    st = np.random.get_state()
    d["np_key_start"], d["np_pos_start"] = st[1].copy(), st[2]
    losses = []
    for e in range(c["epochs"]):
        losses.append([float(x) for x in att.train_step(target_id_list=c["targets"])])
        for k, v in att.netD.state_dict().items():
            d[f"D{e + 1}__{k}"] = v.numpy().copy()
    d["losses"] = np.asarray(losses, dtype=np.float64)
    for k, v in att.netG.state_dict().items():
        d[f"G1__{k}"] = v.numpy().copy()
    g_moved = max(float(np.abs(d[f"G1__{k}"] - d[f"G0__{k}"]).max()) for k in att.netG.state_dict())
    st = np.random.get_state()
    d["np_key_mid"], d["np_pos_mid"] = st[1].copy(), st[2]
    fake = att.generate_fake(target_id_list=c["targets"])
    d["fake"] = np.asarray(fake, dtype=np.float32)
    st = np.random.get_state()
    d["np_key_end"], d["np_pos_end"] = st[1].copy(), st[2]
    d["hp"] = np.asarray([ds.n_users, ds.n_items, ds.config["batch_size"], att.filler_num, att.attack_num, att.ZR_ratio], dtype=np.float64)
    d["selected_ids"] = np.asarray(att.selected_ids, dtype=np.int64)
    d["targets"] = np.asarray(c["targets"], dtype=np.int64)
    print(name, "losses", losses, "generator moved by", g_moved, "fake", fake.shape, float(fake.sum()))
    for k, v in d.items():
        store[f"{name}_{k}"] = v
np.savez_compressed(os.path.join(OUT, "aush_synth.npz"), **store)
