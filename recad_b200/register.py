"""Plug the B200 path into an installed reference (`import recad`), through the reference's own
registries -- no reference file is edited:

    import recad, recad_b200.register
    recad_b200.register.install()            # adds 'lightgcn_b200', 'mf_b200', 'ncf_b200', attacker 'aush_b200', datasets 'implicit_b200' / 'explicit_b200'
    recad_b200.register.install(override=True)   # ALSO rebinds 'lightgcn'/'mf'/'ncf'/'aush'/'implicit'/'explicit' and the evaluator,
                                                 # so an unmodified `recad_runner ...` runs on the CUDA kernels

Registries touched: recad.model.factories['victim'] / ['attacker'] (recad/model/__init__.py:3-18),
recad.default.MODEL['victim'] / ['attacker'] (recad/default.py:104-168), recad.dataset.factories
(recad/dataset/__init__.py:13), and with override=True `Normal.normal_evaluate` /
`Defense.normal_evaluate` (recad/workflow/normal.py:111-160, defense.py:125-174) and the name
`recad.model.attacker.aia.WMFTrainer`, which AIA / Leg-UP look up every attack step to retrain their surrogate
(aia.py:125-207; aushplus.py:11 inherits the method): the CUDA trainer needs no `higher`.
"""
from . import evaluate
from .config import MODEL
from .dataset import ImplicitData
from .victim import factories as victim_factories


def install(override=False):
    import recad  # noqa: F401  (raises ImportError where the reference is absent)
    from recad import dataset as ref_dataset, default as ref_default, model as ref_model, workflow as ref_workflow

    for name, cls in victim_factories.items():
        for key in ([f"{name}_b200", name] if override else [f"{name}_b200"]):
            ref_model.factories["victim"][key] = cls
            ref_default.MODEL["victim"].setdefault(key, dict(MODEL["victim"][name]))
    from .attacker import Aush
    for key in (["aush_b200", "aush"] if override else ["aush_b200"]):
        ref_model.factories["attacker"][key] = Aush
        ref_default.MODEL["attacker"].setdefault(key, dict(MODEL["attacker"]["aush"]))
    from .explicit import ExplicitData
    ref_dataset.factories["implicit_b200"] = ImplicitData
    ref_dataset.factories["explicit_b200"] = ExplicitData
    if override:
        ref_dataset.factories["implicit"] = ImplicitData
        ref_dataset.factories["explicit"] = ExplicitData

        def _normal_evaluate(self, model, model_fake, dataset, target_id_list, topks):
            return evaluate.normal_evaluate(model, model_fake, dataset, target_id_list, topks)
        ref_workflow.Normal.normal_evaluate = _normal_evaluate
        ref_workflow.Defense.normal_evaluate = _normal_evaluate
        from recad.model.attacker import aia as ref_aia
        from .surrogate import WMFTrainer
        ref_aia.WMFTrainer = WMFTrainer
    return True
