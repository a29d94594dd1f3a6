"""Run an UNMODIFIED reference script with its evaluation hot path served by this package.

    python -m mp_reid_b200.dropin /path/to/mp-reid/test_uniprompt.py --config_file configs/ours/cctv_ir_cctv_rgb.yml

The reference imports ``from utils.metrics import R1_mAP_eval`` (processor/processor.py:7,
processor/processor_uniprompt_stage2.py:7) and ``from utils.reranking import re_ranking``
(utils/metrics.py:4).  Its ``utils`` package also holds logger / meter / iotools, so the package is
NOT shadowed: only the two module names are pre-seeded in ``sys.modules`` and everything else keeps
loading from the reference tree.

``MPREID_DROPIN_LOSSES=1`` additionally serves the two training-side distance workloads of SURVEY 8f:
``loss.triplet_loss`` (TripletLoss, loss/make_loss.py:7) and ``loss.supcontrast`` (SupConLoss,
processor/processor_uniprompt_stage1.py:9) resolve to mp_reid_b200.triplet / mp_reid_b200.supcon.
"""
from __future__ import annotations

import os
import runpy
import sys


def install() -> None:
    from . import metrics, reranking
    sys.modules["utils.metrics"] = metrics
    sys.modules["utils.reranking"] = reranking
    pkg = sys.modules.get("utils")
    if pkg is not None:
        pkg.metrics = metrics
        pkg.reranking = reranking
    if os.environ.get("MPREID_DROPIN_LOSSES", "0") == "1":
        from . import supcon, triplet
        sys.modules["loss.triplet_loss"] = triplet
        sys.modules["loss.supcontrast"] = supcon


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m mp_reid_b200.dropin <reference script.py> [script args...]")
    script = os.path.abspath(argv[0])
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))  # the reference scripts import their siblings by name
    install()
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
