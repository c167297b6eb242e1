"""`from utils import ...` of main_for_seq_leave_earlystop_SegMM.py:6-7 -> the B200 path (see ../README.md)."""
from segmminterest_b200.dataset import (BaseReaderSeq_SegMM_sampled, DataCollator, DataLoader, FrameDatasetSeq_SegMM,  # noqa: F401
                                        FrameDatasetSeq_SegMM_sampled)
from segmminterest_b200.reader import BaseReaderSeq_SegMM  # noqa: F401
