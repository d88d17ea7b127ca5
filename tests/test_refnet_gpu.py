"""GPU parity of the engine-backed ReferenceNet (hallo_b200/refnet.py) against the golden fixture
of the UNMODIFIED reference UNet2D + write-mode ReferenceAttentionControl, and against the oracle port on fresh inputs;
plus the hand-off: writer (our ReferenceNet) -> reader.update -> denoising UNet banks."""
import os

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-2          # fp16 kernels vs the fp32 reference (north_star); banks are LayerNorm outputs: much tighter in practice


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def refnet():
    from hallo_b200.models.unet_2d_condition import UNet2DConditionModel
    from hallo_b200.spec import SD15_UNET_CONFIG, UNetConfig
    from hallo_b200.synth import host_threads, synth_state_dict_2d
    dev = _dev()
    torch.set_num_threads(host_threads())
    sd = synth_state_dict_2d(UNetConfig())
    m = UNet2DConditionModel.from_config(SD15_UNET_CONFIG)
    m.load_state_dict(sd, strict=True)
    return m.to(device=dev, dtype=torch.float16), sd


def test_refnet_matches_reference_golden(refnet):
    from hallo_b200.spec import UNetConfig, reader_bank_order
    from hallo_b200.synth import synth_refnet_inputs
    m, _ = refnet
    dev = _dev()
    fx = torch.load(os.path.join(GOLD, "refnet_h8.pt"), weights_only=False)
    inp = synth_refnet_inputs(UNetConfig(), fx["case"]["h"], fx["case"]["h"], seed=fx["case"]["seed"])
    out = m(inp["sample"].to(dev, torch.float16), torch.tensor(inp["timestep"]),
            encoder_hidden_states=inp["encoder_hidden_states"].to(dev, torch.float16), return_dict=False)[0]
    torch.cuda.synchronize()
    err = rel_l2(out, fx["out"].float())
    print(f"refnet h8: last-up-block features rel L2 vs reference = {err:.3e}")
    assert err < TOL
    assert list(m.banks) == [n for n, _ in reader_bank_order(UNetConfig())] == fx["bank_order"]
    for n, b in fx["banks"].items():
        e = rel_l2(m.banks[n], b.float())
        print(f"  bank {n}: rel L2 = {e:.3e}")
        assert e < TOL
    for n, st in fx["bank_stats"].items():
        assert list(m.banks[n].shape) == st["shape"]
        assert abs(float(m.banks[n].float().std()) - st["std"]) < 2e-2 * st["std"], n


@pytest.mark.parametrize("h", [16, 32])
def test_refnet_matches_oracle_port(refnet, h):
    """Fresh inputs: every one of the 16 banks and the output against the fp32 oracle port run on the host."""
    from hallo_b200.spec import UNetConfig
    from hallo_b200.synth import synth_refnet_inputs
    from oracle import port
    m, sd = refnet
    dev = _dev()
    cfg = UNetConfig()
    inp = synth_refnet_inputs(cfg, h, h, seed=100 + h)
    ref_out, ref_banks = port.reference_net_forward(sd, cfg, inp)
    out = m(inp["sample"].to(dev, torch.float16), 0, encoder_hidden_states=inp["encoder_hidden_states"].to(dev, torch.float16),
            return_dict=False)[0]
    torch.cuda.synchronize()
    worst = max(rel_l2(m.banks[n], ref_banks[n]) for n in ref_banks)
    err = rel_l2(out, ref_out)
    print(f"refnet h{h}: output rel L2 = {err:.3e}, worst bank rel L2 = {worst:.3e}")
    assert err < TOL and worst < TOL


def test_writer_to_reader_hand_off(refnet):
    """ReferenceAttentionControl(write) on the engine-backed ReferenceNet -> reader.update(writer) -> the denoising
    UNet's banks: fp16 copies (Q4) of the ReferenceNet's norm1 outputs, keyed by the reference's pairing order."""
    from hallo_b200.models.mutual_self_attention import ReferenceAttentionControl
    from hallo_b200.models.unet_3d import UNet3DConditionModel
    from hallo_b200.spec import HALLO_UNET_KWARGS, SD15_UNET_CONFIG, UNetConfig, reader_bank_order
    from hallo_b200.synth import synth_refnet_inputs
    m, _ = refnet
    dev = _dev()
    unet3d = UNet3DConditionModel.from_config(SD15_UNET_CONFIG, **HALLO_UNET_KWARGS)       # stays on the host: only banks
    writer = ReferenceAttentionControl(m, do_classifier_free_guidance=True, mode="write", batch_size=1, fusion_blocks="full")
    reader = ReferenceAttentionControl(unet3d, do_classifier_free_guidance=True, mode="read", batch_size=1, fusion_blocks="full")
    inp = synth_refnet_inputs(UNetConfig(), 16, 16, seed=5)
    m(inp["sample"].to(dev, torch.float16), 0, encoder_hidden_states=inp["encoder_hidden_states"].to(dev, torch.float16),
      return_dict=False)
    reader.update(writer)
    assert list(unet3d._banks) == [n for n, _ in reader_bank_order(UNetConfig())]
    for n, b in unet3d._banks.items():
        assert b.dtype == torch.float16 and torch.equal(b, m.banks[n].to(torch.float16)) and b.data_ptr() != m.banks[n].data_ptr()
    writer.clear()
    reader.clear()
    assert not m.banks and not unet3d._banks
