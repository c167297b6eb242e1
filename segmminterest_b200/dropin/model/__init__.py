"""`from model import ...` of main_for_seq_leave_earlystop_SegMM.py:5 -> the B200 path (see ../README.md)."""
from segmminterest_b200.evaluation import TOP_K_leave, TOP_K_leave_mask, main_eval_batch  # noqa: F401
from segmminterest_b200.model import MultiScaleTemporalDetrLeaveFocal, QueryBasedDecoder, SegFormerX  # noqa: F401
