"""Device time of one NCF epoch on the ml1m (default) or yelp shape (samples resident), per tower precision; the environment
selects the captured graph (RECAD_NCF_GRAPH) and the lazy embedding Adam (RECAD_NCF_LAZY_ADAM).
    python tools/ncf_epoch_time.py [yelp] [tf32x3|fp32 ...]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recad_b200 import dataset, model, synthetic
DEV = torch.device("cuda:0")
YELP = "yelp" in sys.argv[1:]
PRECS = [a for a in sys.argv[1:] if a in ("tf32x3", "fp32")] or ["tf32x3", "fp32"]
tr, va, te = synthetic.make_splits(synthetic.YELP if YELP else synthetic.ML1M, seed=0)
data = dataset.from_config("implicit", "yelp" if YELP else "ml1m", train_dict=tr, valid_dict=va, test_dict=te, need_graph=False, graph_edges="train",
                           sample="pointwise", device=DEV)
np.random.seed(0)
samples = data.epoch_samples(DEV)
data.epoch_samples = lambda device=None: samples
for prec in PRECS:
    torch.manual_seed(0)
    m = model.from_config("victim", "ncf", tower_precision=prec, device=DEV).I(dataset=data)
    m.train_step(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        loss = m.train_step()[0]
    b.record(); torch.cuda.synchronize()
    n = int(samples[0].shape[0])
    print(json.dumps({"shape": "yelp" if YELP else "ml1m", "tower": prec, "lazy_adam": os.environ.get("RECAD_NCF_LAZY_ADAM", "auto"), "graph": os.environ.get("RECAD_NCF_GRAPH", "1"), "epoch_ms": round(a.elapsed_time(b) / 3, 1),
                      "batches": (n + 1023) // 1024, "us_per_batch": round(a.elapsed_time(b) / 3 / ((n + 1023) // 1024) * 1e3, 1), "loss": loss}))
