"""Mirror of /root/reference/train_constants.py:6-28 (values are part of the drop-in contract)."""
import math

SEQ_LEN_IN = 64
TIME_POOL = [2, 2, 2]
SEQ_LEN_OUT = int(SEQ_LEN_IN // math.prod(TIME_POOL))
BATCH_SIZE = 128
NUM_WORKERS = 4

TIME_MASK_W = 8
FREQ_MASK_W = 8
MASKS_PER_EX = 2

SAMPLE_RATE = 44_100
HOP_LENGTH = 2048 // 2
FPS_ORIG = int(SAMPLE_RATE / HOP_LENGTH)
FPS_OUT = FPS_ORIG // math.prod(TIME_POOL)

N_MELS = 40
CONV_DEPTH = 16
GRU1_UNITS = 16
GRU2_UNITS = 8
DENSE1_UNITS = 8
