"""The registry hook plugs the B200 classes into an installed reference without editing it.
Needs the live reference checkout; skipped on the GPU box."""
import sys

import pytest

pytestmark = pytest.mark.reference


def test_install_registers_victims_dataset_and_evaluator(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                      # the reference resolves ./data and ./generated against cwd
    sys.path.insert(0, "/root/reference")
    try:
        import recad
        import recad_b200.register as reg
        from recad_b200 import dataset, evaluate, victim
        orig_eval = recad.workflow.Normal.normal_evaluate
        assert reg.install() is True
        for name, cls in victim.factories.items():
            assert recad.model.factories["victim"][f"{name}_b200"] is cls
            assert set(recad.default.MODEL["victim"][name]) <= set(recad.default.MODEL["victim"][f"{name}_b200"])
        assert recad.dataset.factories["implicit_b200"] is dataset.ImplicitData
        assert recad.model.factories["victim"]["lightgcn"] is not victim.LightGCN          # not overridden yet
        lazy = recad.model.from_config("victim", "mf_b200", embedding_size=64)              # the reference's own factory
        assert isinstance(lazy, victim.MF) and lazy.model_name == "mf" and not lazy._is_instantiate
        assert lazy._init_config["embedding_size"] == 64 and lazy._init_config["factor_num"] == 3
        reg.install(override=True)
        assert recad.model.factories["victim"]["lightgcn"] is victim.LightGCN
        assert recad.dataset.factories["implicit"] is dataset.ImplicitData
        assert recad.workflow.Normal.normal_evaluate is not orig_eval
        # undo for other tests in the same process
        recad.workflow.Normal.normal_evaluate = orig_eval
        recad.workflow.Defense.normal_evaluate = orig_eval
        recad.model.factories["victim"].update(lightgcn=recad.model.victim.LightGCN, mf=recad.model.victim.MF, ncf=recad.model.victim.NCF)
        recad.dataset.factories["implicit"] = recad.dataset.implicit.ImplicitData
    finally:
        sys.path.remove("/root/reference")


def test_explicit_dataset_equals_the_reference(tmp_path, monkeypatch):
    """recad_b200.explicit.ExplicitData against the live reference class on the dev CSVs: shapes, rating matrix, remap,
    partial_sample (same draws on np.random) and the batch stream."""
    import shutil
    import numpy as np
    import torch
    monkeypatch.chdir(tmp_path)
    shutil.copytree("/root/reference/data/dev", tmp_path / "data" / "dev")
    sys.path.insert(0, "/root/reference")
    try:
        import recad
        from recad_b200 import explicit
        cpu = torch.device("cpu")
        for kw in ({}, {"remap_enable": True}):
            r = recad.dataset.from_config("explicit", "dev", device=cpu, if_cache=False, **kw)
            m = explicit.ExplicitData.from_config("dev", device=cpu, if_cache=False, **kw)
            assert (r.n_users, r.n_items, r.train_size) == (m.n_users, m.n_items, m.train_size)
            assert np.array_equal(r.train_mat, m.train_mat) and np.array_equal(r.train_dict, m.train_dict)
            assert sorted(r.info_describe()) == sorted(m.info_describe())
        np.random.seed(3); rp = r.partial_sample(user_ratio=0.3); s1 = np.random.get_state()
        np.random.seed(3); mp = m.partial_sample(user_ratio=0.3); s2 = np.random.get_state()
        assert rp.n_users == mp.n_users and np.array_equal(rp.train_mat, mp.train_mat) and np.array_equal(rp.test_dict, mp.test_dict)
        assert s1[2] == s2[2] and np.array_equal(s1[1], s2[1])
        flt = lambda train_mat: np.where((train_mat > 0).sum(1) >= 1)[0]
        np.random.seed(4); rb = [(d["users"].numpy(), d["users_mat"].numpy()) for d in rp.generate_batch(user_filter=flt)]
        np.random.seed(4); mb = [(d["users"].numpy(), d["users_mat"].numpy()) for d in mp.generate_batch(user_filter=flt)]
        assert len(rb) == len(mb) and all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(rb, mb))
    finally:
        sys.path.remove("/root/reference")
