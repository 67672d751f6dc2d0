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
