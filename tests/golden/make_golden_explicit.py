"""The explicit `dev` arrays exactly as the reference's ExplicitData loads them (recad/dataset/explicit.py:41-100):
float [n, 3] rows (user, item, rating) for train / valid / test.  They let a test on a box without the reference's data
directory build the SAME attack dataset through the reference's own in-memory entry point
(`dataset.from_config("explicit", "dev", train_dict=..., valid_dict=..., test_dict=...)`, explicit.py:47-52).

    cd <scratch dir with data/dev copied from /root/reference/data/dev>
    PYTHONPATH=/root/reference python /root/repo/tests/golden/make_golden_explicit.py
"""
import os

import numpy as np
import torch

import recad

OUT = os.path.dirname(os.path.abspath(__file__))
ds = recad.dataset.from_config("explicit", "dev", device=torch.device("cpu"), if_cache=False, remap_enable=False)
np.savez_compressed(os.path.join(OUT, "dev_explicit.npz"), train=np.asarray(ds.train_dict), valid=np.asarray(ds.valid_dict),
                    test=np.asarray(ds.test_dict), n_users=ds.n_users, n_items=ds.n_items)
print(ds.n_users, ds.n_items, ds.train_dict.shape, ds.valid_dict.shape, ds.test_dict.shape, ds.train_dict[:3])
