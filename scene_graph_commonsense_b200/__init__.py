"""hiercom-b200: B200-native implementation of HIERCOM's relation-prediction hot path.

Public surface mirrors the reference's modules for this path:
  model.BayesianRelationClassifier / FlatRelationClassifier / BayesianHead    (reference model.py)
  evaluator.Evaluator / Evaluator_Top3                                        (reference evaluator.py)
  pipeline.RelationPipeline  - the batched entry point over whole images (pairs -> counters)
  sgb.BayesHead / HierarchPostProcessor                                       (Scene-Graph-Benchmark plug-in)
  losses.RelationLoss        - training-side losses of train_utils.train_one_direction on the fused head, with backward
"""
from . import tables  # noqa: F401

__all__ = ["tables", "ops", "model", "evaluator", "pipeline", "synthetic", "dist", "sgb", "losses"]
