"""The drop-in proven through the reference's OWN workflow: the unmodified reference package (installed under
baseline/_ref, it travels to the GPU box) builds a "no defense" workflow with its own factories, its own explicit
attack dataset and its own RandomAttacker; `recad_b200.register.install(override=True)` has rebound the victim classes,
the implicit dataset and the evaluator to the CUDA path; `Normal.execute()` (recad/workflow/normal.py:162-225) then
reproduces the table the live all-reference CPU run printed (tests/golden/meta.json, made by tests/golden/make_golden.py).
"""
import os
import random
import sys

import numpy as np
import pytest
import torch

from . import util

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


@pytest.fixture()
def ref_recad(tmp_path, monkeypatch):
    if not os.path.isdir(os.path.join(REF, "recad")):
        pytest.skip("baseline/_ref (pip install --target of the reference) is absent")
    monkeypatch.chdir(tmp_path)                       # the reference resolves ./data and ./generated against cwd
    sys.path.insert(0, REF)
    try:
        import recad
        import recad_b200.register as reg
        from recad.model.attacker import aia as ref_aia
        saved = (dict(recad.model.factories["victim"]), dict(recad.dataset.factories), recad.workflow.Normal.normal_evaluate,
                 recad.workflow.Defense.normal_evaluate, dict(recad.model.factories["attacker"]))
        saved_wmf = ref_aia.WMFTrainer
        reg.install(override=True)
        yield recad
        ref_aia.WMFTrainer = saved_wmf
        recad.model.factories["victim"].clear(); recad.model.factories["victim"].update(saved[0])
        recad.model.factories["attacker"].clear(); recad.model.factories["attacker"].update(saved[4])
        recad.dataset.factories.clear(); recad.dataset.factories.update(saved[1])
        recad.workflow.Normal.normal_evaluate, recad.workflow.Defense.normal_evaluate = saved[2], saved[3]
    finally:
        sys.path.remove(REF)


@pytest.mark.parametrize("victim,kw,sample", [("lightgcn", {"latent_dim_rec": 64}, "pairwise"), ("mf", {"embedding_size": 64}, "pointwise")])
def test_reference_execute_runs_on_the_cuda_path(ref_recad, victim, kw, sample):
    recad = ref_recad
    from recad_b200 import dataset as b_dataset, victim as b_victim
    dev = torch.device("cuda:0")
    z = util.load(f"workflow_{victim}_dev.npz")
    gold = util.meta()[f"workflow_{victim}_dev"]
    ex = util.load("dev_explicit.npz")
    tr, va, te = util.dicts("dev")
    random.seed(2023); np.random.seed(2023); torch.manual_seed(2023)
    cfg = {
        # the reference's own factories (recad/dataset/__init__.py:13, recad/model/__init__.py:3-21)
        "victim_data": recad.dataset.from_config("implicit", "dev", need_graph=victim == "lightgcn", sample=sample, device=dev,
                                                 train_dict=tr, valid_dict=va, test_dict=te),
        "attack_data": recad.dataset.explicit.ExplicitData.from_config("dev", device=torch.device("cpu"), download=False, train_dict=ex["train"].copy(),
                                                 valid_dict=ex["valid"].copy(), test_dict=ex["test"].copy()).partial_sample(user_ratio=0.2),
        "victim": recad.model.from_config("victim", victim, device=dev, **kw),
        "attacker": recad.model.from_config("attacker", "random", filler_num=36, device=torch.device("cpu")),
        "rec_epoch": gold["rec_epoch"], "attack_epoch": 1, "device": dev,
    }
    assert isinstance(cfg["victim_data"], b_dataset.ImplicitData) and isinstance(cfg["victim"], b_victim.factories[victim])
    assert type(cfg["attacker"]).__module__.startswith("recad.") and type(cfg["attack_data"]).__module__.startswith("recad.")
    wf = recad.workflow.from_config("no defense", **cfg)
    assert type(wf).__module__ == "recad.workflow.normal"
    # same starting point as the golden run: initial weights and both generator states as recorded there
    init = {k[len("init__"):]: z[k] for k in z.files if k.startswith("init__")}
    sd = wf.victim.state_dict()
    rng_ok = all(np.array_equal(sd[k].cpu().numpy(), v) for k, v in init.items() if k in sd)
    for k, v in init.items():
        if k in sd:
            sd[k].copy_(torch.as_tensor(v))
    np.random.set_state(("MT19937", z["np_key_start"], int(z["np_pos_start"]), 0, 0.0))
    torch.set_rng_state(torch.as_tensor(z["torch_state_start"]))
    seen = {}
    gen_fake = wf.attacker.generate_fake

    def cap_fake(**kwargs):
        seen["fake"] = gen_fake(**kwargs)
        return seen["fake"]
    wf.attacker.generate_fake = cap_fake
    ev = type(wf).normal_evaluate

    def cap_eval(self, *a, **k):
        seen["table"] = ev(self, *a, **k)
        return seen["table"]
    type(wf).normal_evaluate = cap_eval
    try:
        wf.execute()                                  # the reference's own driver, recad/workflow/normal.py:162-225
    finally:
        type(wf).normal_evaluate = ev
    # the reference's attacker drew its profiles from np.random AFTER our sampler consumed the stream during step 1:
    # equal fake profiles = the sampler left the generator exactly where the reference's Python loop leaves it
    fake = np.zeros(tuple(z["fake_shape"]), dtype=np.float64)
    fake[z["fake_rows"], z["fake_cols"]] = z["fake_vals"]
    assert np.array_equal(np.asarray(seen["fake"], dtype=np.float64), fake)
    final = {k[len("final__"):]: z[k] for k in z.files if k.startswith("final__")}
    sd = wf.victim.state_dict()
    for k, v in final.items():
        if k in sd and k != "mean":
            assert np.allclose(sd[k].cpu().numpy(), v, rtol=1e-4, atol=2e-6), k
    n_eval = 310
    for k, v in gold["table"].items():
        if "after attack" in k or k == "pred_shift":
            if not rng_ok:
                continue                               # the attacked model's init needs the same torch CPU stream
            tol = dict(rtol=1e-2, atol=2e-6) if k == "pred_shift" else dict(rtol=2e-3, atol=1.5 / n_eval)
            assert np.isclose(seen["table"][k], v, **tol), (k, seen["table"][k], v)
        else:
            assert np.isclose(seen["table"][k], v, rtol=1e-4, atol=1.0 / n_eval), (k, seen["table"][k], v)


def test_reference_attacker_retrains_its_surrogate_on_the_cuda_path(ref_recad):
    """AIA.get_sur_predictions (recad/model/attacker/aia.py:125-207, unmodified; Leg-UP inherits it) builds whatever
    `WMFTrainer` its module names.  After install(override=True) that is the CUDA trainer: the reference method returns
    predictions whose gradient reaches the generator's fake profiles without `higher` being installed."""
    import types
    from recad.model.attacker import aia as ref_aia
    from recad_b200 import surrogate
    assert ref_aia.WMFTrainer is surrogate.WMFTrainer
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    U, I, F = 90, 60, 10
    train = ((rng.random((U, I)) < 0.2) * rng.integers(1, 6, (U, I))).astype(np.float32)
    me = types.SimpleNamespace(train_array=train, device=dev, n_users=U, n_items=I, attack_num=F, surrogate="WMF",
                               config={"hidden_dim_s": 16, "lr_s": 1e-2, "weight_decay_s": 1e-5, "batch_size_s": 16,
                                       "weight_pos_s": 1.0, "weight_neg_s": 0.0, "epoch_s": 4, "unroll_steps_s": 1})
    fake = torch.tensor(rng.random((F, I)).astype(np.float32) * 5, device=dev, requires_grad=True)
    torch.manual_seed(0); np.random.seed(0)
    pred = ref_aia.AIA.get_sur_predictions(me, fake)
    assert pred.shape == (U + F, I) and pred.is_cuda
    pred[:U, 3].sum().backward()                       # an attack-style loss on a target item's scores of the genuine users
    assert fake.grad is not None and float(fake.grad.abs().max()) > 0
    # same call on the oracle (CPU): same predictions, same gradient to the fake profiles
    from oracle import wmf as owmf
    torch.manual_seed(0); np.random.seed(0)
    fake_c = fake.detach().cpu().clone().requires_grad_(True)
    data_c = torch.cat([torch.from_numpy(train), fake_c], 0)
    pred_c, _, _ = owmf.fit_adv(data_c, 4, 1)
    pred_c[:U, 3].sum().backward()
    assert np.abs(pred.detach().cpu().numpy() - pred_c.detach().numpy()).max() <= 1e-4 * float(pred_c.abs().max()) + 1e-6
    assert np.abs(fake.grad.cpu().numpy() - fake_c.grad.numpy()).max() <= 2e-3 * float(fake_c.grad.abs().max())


def test_reference_dataset_and_factory_drive_the_cuda_aush(ref_recad):
    """`recad.model.from_config("attacker", "aush")` (the reference's factory) resolves to the CUDA attacker after
    install(override=True); fed by the reference's own ExplicitData batch generator (recad/dataset/explicit.py:166-188) it
    reproduces the golden run of the reference's Aush on the same dataset: epoch losses, generator state, fake profiles."""
    recad = ref_recad
    from recad_b200 import attacker as b_attacker
    dev = torch.device("cuda:0")
    mat, G, D, kw, batch, targets, z = util.aush_case("a")
    tr, te = z["train"].astype(np.float64), z["test"].astype(np.float64)
    random.seed(2023); np.random.seed(2023); torch.manual_seed(2023)
    ds = recad.dataset.explicit.ExplicitData.from_config("dev", device=dev, download=False, if_cache=False, remap_enable=False, train_dict=tr.copy(),
                                   valid_dict=te.copy(), test_dict=te.copy())
    assert type(ds).__module__.startswith("recad.")
    att = recad.model.from_config("attacker", "aush", device=dev).I(dataset=ds)
    assert isinstance(att, b_attacker.Aush)
    for k, v in att.netG_state().items():                      # same torch.manual_seed -> the reference's initial networks
        assert np.array_equal(v.cpu().numpy(), G[k]), k
    for k, v in att.netD_state().items():
        assert np.array_equal(v.cpu().numpy(), D[k]), k
    st = np.random.get_state()
    assert np.array_equal(st[1], z["a_np_key_start"]) and st[2] == int(z["a_np_pos_start"])
    att = att.to(dev)
    for e, gold in enumerate(z["a_losses"]):
        loss = att.train_step(target_id_list=targets, input_describe={}, progress_bar=None)
        assert np.allclose(loss, gold, rtol=1e-4, atol=0), (e, loss, gold)
    fake = att.generate_fake(target_id_list=targets)
    st = np.random.get_state()
    assert np.array_equal(st[1], z["a_np_key_end"]) and st[2] == int(z["a_np_pos_end"])
    assert np.mean(fake != z["a_fake"]) <= 2.0 / fake.size
