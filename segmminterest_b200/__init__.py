"""segmminterest_b200 -- B200-native MMinterest training step (hezy18/SegMMInterest hot path).

Public surface (mirrors MMinterest/models + the loop in main_for_seq_leave_earlystop_SegMM.py):
    SegFormerX, MultiScaleTemporalDetrLeaveFocal, QueryBasedDecoder, build_model   (model.py)
    DeviceGather, TrainStep                                                        (train.py)
Everything computes through libmmi_b200.so (include/mmi_b200.h); no CPU fallback.
"""
__version__ = "0.1.0"

from .model import (InteractionAggregation, MultiScaleTemporalDetrLeaveFocal, QueryBasedDecoder, SegFormerX,  # noqa: E402,F401
                    build_model)
from .train import DeviceGather, TrainStep  # noqa: E402,F401
