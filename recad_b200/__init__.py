"""recad_b200: the RecAD victim-model hot path (LightGCN / MF / NCF train step, full-ranking
evaluation, normalised-adjacency construction with in-place fake-user injection) as hand-written
sm_100a CUDA kernels behind the reference's own Python plugin interface.

    from recad_b200 import dataset, model, workflow
    data = dataset.from_config("implicit", "ml1m", need_graph=True)
    victim = model.from_config("victim", "lightgcn", latent_dim_rec=64).I(dataset=data)
    victim.train_step()

Mirrors `recad.dataset.from_config`, `recad.model.from_config`, `recad.workflow.from_config`
(recad/dataset/__init__.py:13-17, recad/model/__init__.py:20-21, recad/workflow/__init__.py:4-7).
`recad_b200.register.install()` plugs the same classes into an installed reference.
"""
__version__ = "0.1.0"
