"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's prior decoding.

  mv2mvs                     test_LD_37.py:83-105 (= train_LD_37.py:136-158, test_LD_22_FPS.py:100-122)
  modify_mv_for_end_frames   test_LD_37.py:209-234
  generate_input_index       test_LD_37.py:13-16

Pinned by tests/golden/priors_golden.npz, produced by oracle/make_golden.py
from the reference's own functions (extracted from test_LD_37.py with `ast`,
because the script itself imports a module that does not exist).
"""
import numpy as np


def mv2mvs(mv):
    """mv int [H,W,3] -> fp32 [7,H,W,2]; last dim (x, y). Every step is one IEEE fp32 operation."""
    m = mv.astype(np.float32)
    fx_src, fy_src, rd = m[:, :, 1], m[:, :, 0], m[:, :, 2]  # the reference swaps channels 0 and 1 first
    den = rd * np.float32(-1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        fx = fx_src / den
        fy = fy_src / den
    fx = np.where(np.isnan(fx), np.float32(0), fx).astype(np.float32)
    fy = np.where(np.isnan(fy), np.float32(0), fy).astype(np.float32)
    out = np.zeros((7,) + mv.shape[:2] + (2,), np.float32)
    base = np.stack([fx, fy], axis=-1)
    with np.errstate(invalid="ignore"):
        for f, s in ((2, None), (1, 2.0), (0, 3.0), (4, -1.0), (5, -2.0), (6, -3.0)):
            out[f] = base if s is None else base * np.float32(s)
        out = out / np.float32(128.0)
    return out


def mv2mvs_model_layout(mv):
    """[1,7,2,H,W] as the model receives it (unsqueeze + permute(0,1,4,2,3), test_LD_37.py:160-161)."""
    return np.ascontiguousarray(mv2mvs(mv)[None].transpose(0, 1, 4, 2, 3))


def modify_mv_for_end_frames(i, mvs, max_idx):
    """In place on [B,7,2,H,W]; same statement order as the reference."""
    if i == 0:
        mvs[:, 0:3] = 0.0
    if i == 1:
        mvs[:, 0] = mvs[:, 2]
        mvs[:, 1] = mvs[:, 2]
    if i == 2:
        mvs[:, 0] = mvs[:, 1]
    if i == max_idx - 1:
        mvs[:, 4:7] = 0.0
    if i == max_idx - 2:
        mvs[:, 5] = mvs[:, 4]
        mvs[:, 6] = mvs[:, 4]
    if i == max_idx - 3:
        mvs[:, 6] = mvs[:, 5]
    return mvs


def generate_input_index(center_index, frame_number, max_index):
    o = np.arange(frame_number) - frame_number // 2 + center_index
    return np.clip(o, 0, max_index)
