"""Host I/O either side of the path, on the device.

Mirrors the two helpers the reference's inference script wraps around `model(input_im)`
(run_ood_faceGAN_inversion.py:158-174):

    cv2im = cv2.imread(file) / 255.0
    input_im = (torch.stack(img2tensor([cv2im], bgr2rgb=True), dim=0) - 0.5) * 2        # BasicSR img_util.py:10-36
    ...
    img = tensor2img(inversion_im, rgb2bgr=True, min_max=(-1, 1))                        # BasicSR img_util.py:38-94

Same names and argument meaning; the difference is where the bytes are converted: `img2tensor` takes the decoded uint8
frames (cv2's BGR [H,W,3] layout, stacked to [B,H,W,3]) already on the GPU and returns the normalised fp32 batch,
`tensor2img` returns uint8 [B,H,W,3] on the GPU, so the serving loop moves 3 bytes per pixel each way over PCIe instead
of 12.  Results are bit-identical to the reference's CPU arithmetic (tests/test_kernels_gpu.py::test_imgio_*).  CUDA only.
"""
import torch

from . import kernels as K


def img2tensor(imgs, bgr2rgb=True, float32=True, mean=0.5, scale=2.0):
    """uint8 frames [B,H,W,3] or [H,W,3] (CUDA) -> fp32 [B,3,H,W] = (float32(v / 255.0) - mean) * scale.

    `bgr2rgb`, `float32`: as BasicSR's img2tensor (img_util.py:10-36); the division by 255 that the caller does on the
    numpy side (run_ood_faceGAN_inversion.py:158) and the normalisation `(t - 0.5) * 2` (:159) are part of the kernel."""
    if not float32:
        raise NotImplementedError('img2tensor: float32=False (uint8 planes) is not on the hot path')
    if isinstance(imgs, (list, tuple)):
        imgs = torch.stack(list(imgs), 0)
    if imgs.dim() == 3:
        imgs = imgs.unsqueeze(0)
    return K.img2tensor_u8(imgs.contiguous(), swap_rb=bgr2rgb, sub=mean, mul=scale)


def tensor2img(tensor, rgb2bgr=True, out_type=torch.uint8, min_max=(0, 1)):
    """fp32 [B,3,H,W] or [3,H,W] (CUDA) -> uint8 [B,H,W,3] (BGR if rgb2bgr), BasicSR tensor2img (img_util.py:38-94) per image:
    clamp to min_max, (t - min) / (max - min), (* 255.0).round().  A 4-D input is NOT tiled into a grid (the reference's
    make_grid branch, :70-74, is a visualisation aid): every image is converted on its own, as the script's per-file loop does."""
    if out_type not in (torch.uint8, 'uint8'):
        raise NotImplementedError('tensor2img: only the uint8 output of the inference script is implemented')
    if tensor.dim() == 3:
        tensor = tensor.unsqueeze(0)
    if tensor.dim() != 4 or tensor.shape[1] != 3:
        raise ValueError(f'tensor2img: expected [B,3,H,W] or [3,H,W], got {tuple(tensor.shape)}')
    lo, hi = float(min_max[0]), float(min_max[1])
    return K.tensor2img_u8(tensor.detach().float().contiguous(), swap_rb=rgb2bgr, lo=lo, hi=hi)


class ByteServing(torch.nn.Module):
    """`net` between the two converters: uint8 BGR frames in, uint8 BGR frames out (what `graphs.PipelinedForward` captures
    for the byte-format serving loop).  net(x) must return the image first, as ood_faceGAN_e4e.forward does."""

    def __init__(self, net, min_max=(-1, 1)):
        super().__init__()
        self.net, self.min_max = net, min_max

    def forward(self, frames):
        out = self.net(img2tensor(frames))
        return tensor2img(out[0] if isinstance(out, (tuple, list)) else out, min_max=self.min_max)
