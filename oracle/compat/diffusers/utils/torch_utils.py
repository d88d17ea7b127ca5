"""apply_freeu: import-only (FreeU is never enabled by the reference, unet_2d_blocks.py:27)."""


def apply_freeu(resolution_idx, hidden_states, res_hidden_states, **freeu_kwargs):
    raise NotImplementedError("FreeU is not enabled in the reference pipeline")
