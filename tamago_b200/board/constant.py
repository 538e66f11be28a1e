"""Board constants of the reference (board/constant.py:4-31) that the host surfaces need."""
PASS = 0
RESIGN = -1
OB_SIZE = 1
GTP_X_COORDINATE = "IABCDEFGHJKLMNOPQRSTUVWXYZ"
