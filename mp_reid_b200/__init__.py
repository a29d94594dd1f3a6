"""mp_reid_b200 — B200-native (sm_100a) evaluation / retrieval hot path of MP-ReID.

Reference-facing modules: ``mp_reid_b200.metrics`` (utils/metrics.py) and
``mp_reid_b200.reranking`` (utils/reranking.py); ``mp_reid_b200.dropin`` makes the unmodified
reference scripts resolve ``utils.metrics`` / ``utils.reranking`` to them.
"""
__version__ = "0.1.0"
