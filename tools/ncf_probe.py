"""One NCF batch in tensor-core vs exact mode (run twice with/without RECAD_NCF_EXACT=1 and diff the dumps)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recad_b200 import model
DEV = torch.device("cuda:0")
class Stub:
    def __init__(s, U, I, batches):
        s.U, s.I, s.batches = U, I, batches
        s.config = {"pointwise_batch_size": 1024, "device": DEV}
    def info_describe(s): return {"n_users": s.U, "n_items": s.I, "train_dict": None}
    def generate_batch(s):
        for u, i, y in s.batches:
            yield {"users": torch.as_tensor(u), "items": torch.as_tensor(i), "labels": torch.as_tensor(y)}
g = np.random.default_rng(0)
U, I, n = 512, 511, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
b = [(g.integers(0, U, n), g.integers(0, I, n), g.integers(0, 2, n))]
torch.manual_seed(1)
f, L = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (32, 5)
m = model.from_config("victim", "ncf", factor_num=f, num_layers=L, device=DEV).I(dataset=Stub(U, I, b))
loss = m.train_step()[0]
torch.cuda.synchronize()
np.savez(sys.argv[1], loss=loss, g=m.g.cpu().numpy(), flat=m.flat.cpu().numpy())
print("loss", loss)
