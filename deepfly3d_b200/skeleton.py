"""Constants of the 38-joint fly skeleton used on the hot path (values from the reference's
df3d/skeleton_fly.py:6-55,190 and df3d/config.py:15-21; only what packing / procrustes need)."""
import numpy as np

NUM_CAMERAS = 7
NUM_JOINTS = 38
NUM_PREDICT = NUM_JOINTS // 2          # 19 maps per image (config.py:36)
HEATMAP_SHAPE = (64, 128)              # config.py:18

# joint type per index inside one 19-joint half: three 5-joint legs, antenna, three stripes
BODY_COXA, COXA_FEMUR, FEMUR_TIBIA, TIBIA_TARSUS, TARSUS_TIP, ANTENNA, STRIPE = range(7)
HALF_TYPES = [BODY_COXA, COXA_FEMUR, FEMUR_TIBIA, TIBIA_TARSUS, TARSUS_TIP] * 3 + [ANTENNA, STRIPE, STRIPE, STRIPE]
TRACKED_POINTS = HALF_TYPES + HALF_TYPES

# joints used for the rigid alignment (procrustes.py:55: BODY_COXA and COXA_FEMUR)
ALIGN_IDX = np.array([j for j, t in enumerate(HALF_TYPES) if t in (BODY_COXA, COXA_FEMUR)])
N_LEGS = 3
