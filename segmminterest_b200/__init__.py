"""segmminterest_b200 -- B200-native MMinterest training step (hezy18/SegMMInterest hot path).

Public surface (mirrors MMinterest/models + the loop in main_for_seq_leave_earlystop_SegMM.py):
    SegFormerX, MultiScaleTemporalDetrLeaveFocal, QueryBasedDecoder, build_model   (model.py)
    DeviceGather, TrainStep                                                        (train.py)
    InferenceScorer (forward-only batched scoring, BASELINE config 5)              (inference.py)
    SegmentIndex, DeviceFrameLoader (dataset + DataLoader + DataCollator drop-in)  (index.py, loader.py)
    load_feature_table (SegMM_feat_memmap.dat, float32 or float64 -> resident table) (table.py)
    main_eval_batch, prob_auc_batch (device ProbAUC / per-row validation metrics)  (evaluation.py)
Everything computes through libmmi_b200.so (include/mmi_b200.h); no CPU fallback.
"""
__version__ = "0.1.0"

from .model import (InteractionAggregation, MultiScaleTemporalDetrLeaveFocal, QueryBasedDecoder, SegFormerX,  # noqa: E402,F401
                    build_model)
from .train import DeviceGather, TrainStep  # noqa: E402,F401
from .inference import InferenceScorer  # noqa: E402,F401
from .loader import DeviceFrameLoader  # noqa: E402,F401
from .index import SegmentIndex  # noqa: E402,F401
from .table import load_feature_table  # noqa: E402,F401
from .evaluation import TOP_K_leave, TOP_K_leave_mask, main_eval_batch, prob_auc_batch  # noqa: E402,F401
