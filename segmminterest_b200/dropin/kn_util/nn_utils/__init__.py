"""`from kn_util.nn_utils import CheckPointer` (main_for_seq_leave_earlystop_SegMM.py:16,217)."""
from segmminterest_b200.checkpoint import CheckPointer  # noqa: F401
