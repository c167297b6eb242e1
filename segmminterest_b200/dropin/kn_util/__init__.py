"""`kn_util` as far as the driver needs it (main_for_seq_leave_earlystop_SegMM.py:16)."""
