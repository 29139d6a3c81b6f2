"""SURVEY.md Appendix C, typed in as literals: the known-answer vectors the survey lists for this path.
The neutral-pose EE position is a real-MuJoCo number (scripts/execute_pnp.py:38) and the first reward
row is the reference's own pickled `old_reward`; the rest were derived by the survey's restatement."""
import numpy as np

NEUTRAL = np.array([0.0, 0.41, 0.0, -1.85, 0.0, 2.26, 0.79])
FK_NEUTRAL_POS = np.array([1.23843967, 0.0, 0.49740014])
FK_NEUTRAL_QUAT = np.array([0.0, 0.99999735, -0.00230093, 0.0])
FK_NEUTRAL_JP = np.array([[0.0, -0.13559986, 0.0, 0.39252477, 0.0, 0.212, 0.0],
                          [0.63843967, 0.0, 0.63957768, 0.0, 0.10765036, 0.0, 0.0],
                          [0.0, -0.63843967, 0.0, 0.43681665, 0.0, 0.088, 0.0]])
FK_NEUTRAL_JR = np.array([[0.0, 0.0, 0.39860933, 0.0, 0.77175266, 0.0, 0.0],
                          [0.0, 1.0, 0.0, -1.0, 0.0, -1.0, 0.0],
                          [1.0, 0.0, 0.91712082, 0.0, -0.63592282, 0.0, -1.0]])
FK_ZERO_POS = np.array([0.688, 0.0, 1.121])
FK_ZERO_QUAT = np.array([0.0, 0.92387954, 0.38268342, 0.0])

# (target, kwargs, iterations, q, final_pos, pos_error) - cold from neutral
IK_CASES = [
    ((1.415, 0.0, 0.73), {}, 7, [0, 0.4737271, 0, -1.4312119, 0, 2.63849197, 0.79], [1.41469724, 0, 0.73000737], 3.0285e-4),
    ((1.415, 0.0, 1.03), {}, 11, [0, 0.44360306, 0, -1.06209243, 0, 2.85524461, 0.79], [1.41410427, 0, 1.02964457], 9.6367e-4),
    ((1.415, 0.0, 0.43), {}, 8, [0, 0.88977843, 0, -1.37925763, 0, 2.71043638, 0.79], [1.41475575, 0, 0.43009582], 2.6238e-4),
    ((1.23843967, 0.0, 0.49740014), {}, 1, NEUTRAL.tolist(), [1.23843967, 0, 0.49740014], 3.5e-9),
    ((1.33843967, 0.0, 0.49740014), dict(pos_thresh=1e-4, damping=0.05), 10,
     [0, 0.59667011, 0, -1.62267184, 0, 2.46372152, 0.79], [1.33834829, 0, 0.49743646], 9.833e-5),
]

H = 0.7071067811865476
HORIZONTAL_QUAT = (H, -0.7071067811865475, 0.0, 0.0)
# (ag, dg, ee_pos, width, ee_quat wxyz, task_idx, dense value, dense bits, sparse bits); h0 = 0.001, len = 3, threshold 0.05
REWARD_ROWS = [
    ((1.46172, 0.17519, 0.01989), (1, -0.1, 0.3), (1.3847, 0.35938, 0.58267), 0.07983, (0, 1, 0, 0), 0, -0.053, 0xBD591687, 0xBF800000),
    ((1.4, 0, 0.73), (1, -0.1, 0.3), (1.4, 0.02, 0.73), 0.08, (0, 1, 0, 0), 0, -0.023, 0xBCBC6A7F, 0xBF800000),
    ((1.4, 0, 0.73), (1, -0.1, 0.3), (1.41, 0, 0.73), 0.04, HORIZONTAL_QUAT, 0, 6.987, 0x40DF9581, 0xBF800000),
    ((1.0, 0.1, 0.32), (1, -0.1, 0.3), (1.0, 0.1, 0.33), 0.04, (1, 0, 0, 0), 1, 7.1536665, 0x40E4EAD6, 0xBF800000),
    ((1.0, -0.1, 0.32), (1, -0.1, 0.3), (1.0, -0.1, 0.33), 0.04, (0, 1, 0, 0), 2, 16.320333, 0x4182900B, 0x80000000),
    ((1.0, -0.1, 0.32), (1, -0.1, 0.3), (1.0, -0.1, 0.45), 0.08, (0, 1, 0, 0), 0, 9.9469995, 0x411F26E9, 0x80000000),
]


def reward_arrays(dtype=np.float64):
    ag = np.array([r[0] for r in REWARD_ROWS], dtype=dtype)
    dg = np.array([r[1] for r in REWARD_ROWS], dtype=dtype)
    ee = np.array([r[2] for r in REWARD_ROWS], dtype=dtype)
    w = np.array([r[3] for r in REWARD_ROWS], dtype=dtype)
    eq = np.array([r[4] for r in REWARD_ROWS], dtype=dtype)
    idx = np.array([r[5] for r in REWARD_ROWS], dtype=np.int32)
    dense_bits = np.array([r[7] for r in REWARD_ROWS], dtype=np.uint32)
    sparse_bits = np.array([r[8] for r in REWARD_ROWS], dtype=np.uint32)
    return ag, dg, ee, eq, w, idx, dense_bits, sparse_bits
