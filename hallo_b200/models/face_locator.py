"""FaceLocator (hallo/models/face_locator.py:31-118): face-region mask -> 320-channel conditioning added after conv_in
(unet_3d.py:605).  Per-frame 2-D convs on a mask that is IDENTICAL for every frame of a window and every window of a
clip (face_animate.py:338-342 repeats one mask video_length times), so the encoder runs on ONE frame and the result is
expanded (conditioning hoisted out of the window loop).  Same state-dict keys as the reference
(conv_in, blocks.{0..5}, conv_out); plain PyTorch, outside the hot path."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class FaceLocator(nn.Module):
    def __init__(self, conditioning_embedding_channels: int, conditioning_channels: int = 3,
                 block_out_channels=(16, 32, 64, 128)):
        super().__init__()
        self.conv_in = nn.Conv2d(conditioning_channels, block_out_channels[0], 3, padding=1)
        self.blocks = nn.ModuleList([])
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self.blocks.append(nn.Conv2d(cin, cin, 3, padding=1))
            self.blocks.append(nn.Conv2d(cin, cout, 3, padding=1, stride=2))
        self.conv_out = nn.Conv2d(block_out_channels[-1], conditioning_embedding_channels, 3, padding=1)
        nn.init.zeros_(self.conv_out.weight)           # zero_module (face_locator.py:84-91)
        nn.init.zeros_(self.conv_out.bias)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def encode_frame(self, x):
        """(n, 3, H, W) -> (n, C, H/8, W/8)"""
        e = F.silu(self.conv_in(x))
        for blk in self.blocks:
            e = F.silu(blk(e))
        return self.conv_out(e)

    @torch.no_grad()
    def forward(self, conditioning):
        """(bs, 3, f, H, W) -> (bs, C, f, H/8, W/8), the reference's contract.  Frames that are views of one frame
        (stride 0 along f, as the pipeline's repeat produces after our expand) are encoded once."""
        bs, c, f, H, W = conditioning.shape
        if f > 1 and conditioning.stride(2) == 0:
            e = self.encode_frame(conditioning[:, :, 0])
            return e.unsqueeze(2).expand(bs, e.shape[1], f, e.shape[2], e.shape[3])
        x = conditioning.permute(0, 2, 1, 3, 4).reshape(bs * f, c, H, W)
        e = self.encode_frame(x)
        return e.reshape(bs, f, *e.shape[1:]).permute(0, 2, 1, 3, 4)
