"""ORACLE (test infrastructure, never shipped or timed as the product): CPU restatement of the reference's image byte formats.

numpy, exactly the operations of
  * run_ood_faceGAN_inversion.py:158-159 + BasicSR/basicsr/utils/img_util.py:10-36  (frame -> normalised tensor)
  * BasicSR/basicsr/utils/img_util.py:38-94 as called at run_ood_faceGAN_inversion.py:68  (tensor -> frame)
with cv2.cvtColor(BGR2RGB / RGB2BGR) restated as a reversal of the last axis (it is one for 3-channel float / uint8 data).
Pinned against the unmodified reference functions by tests/golden/imgio.pt (tests/golden/make_golden.py).
"""
import numpy as np
import torch


def frame_to_tensor(frame_u8, bgr2rgb=True):
    """uint8 [H,W,3] ndarray -> fp32 tensor [3,H,W] in [-1, 1]."""
    img = frame_u8 / 255.0                                   # float64, run_ood_faceGAN_inversion.py:158
    if bgr2rgb:
        img = img.astype('float32')[..., ::-1]               # img_util.py:24-27
    t = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).float()      # :28-30
    return (t - 0.5) * 2                                     # run_ood_faceGAN_inversion.py:159


def tensor_to_frame(t, rgb2bgr=True, min_max=(-1, 1)):
    """fp32 tensor [3,H,W] (or [1,3,H,W]) -> uint8 [H,W,3] ndarray."""
    t = t.squeeze(0).float().detach().cpu().clone().clamp_(*min_max)                # img_util.py:66
    t = (t - min_max[0]) / (min_max[1] - min_max[0])                                # :67
    img = t.numpy().transpose(1, 2, 0)                                              # :76-77
    if rgb2bgr:
        img = img[..., ::-1]                                                        # :81-82
    img = (img * 255.0).round()                                                     # :89
    return np.ascontiguousarray(img.astype(np.uint8))                               # :90
