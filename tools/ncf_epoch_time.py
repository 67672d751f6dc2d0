"""Device time of one NCF epoch on the ml1m shape (samples resident), per tower precision, with / without the captured graph."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recad_b200 import dataset, model, synthetic
DEV = torch.device("cuda:0")
tr, va, te = synthetic.make_splits(synthetic.ML1M, seed=0)
data = dataset.from_config("implicit", "ml1m", train_dict=tr, valid_dict=va, test_dict=te, need_graph=False, graph_edges="train",
                           sample="pointwise", device=DEV)
np.random.seed(0)
samples = data.epoch_samples(DEV)
data.epoch_samples = lambda device=None: samples
for prec in ("tf32x3", "fp32"):
    torch.manual_seed(0)
    m = model.from_config("victim", "ncf", tower_precision=prec, device=DEV).I(dataset=data)
    m.train_step(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        loss = m.train_step()[0]
    b.record(); torch.cuda.synchronize()
    n = int(samples[0].shape[0])
    print(json.dumps({"tower": prec, "graph": os.environ.get("RECAD_NCF_GRAPH", "1"), "epoch_ms": round(a.elapsed_time(b) / 3, 1),
                      "batches": (n + 1023) // 1024, "us_per_batch": round(a.elapsed_time(b) / 3 / ((n + 1023) // 1024) * 1e3, 1), "loss": loss}))
