"""Host-side kernel PLANS executed without a GPU: hallo_b200/engine.py and hallo_b200/refnet.py run with the C-ABI
wrappers of hallo_b200.ops replaced by the PyTorch stand-ins of tests/cpu_ops.py (test infrastructure), in fp32, and
are compared with the oracle.  This checks everything the host decides -- weight packing, buffer reuse and aliasing,
row offsets of the 18-frame buffers, skip wiring, hoisted window constants, the frame <-> pixel bookkeeping of a
sharded window, the step tail -- before any GPU time is spent; the kernels themselves are the `-m gpu` tests' job."""
import pytest
import torch

import cpu_ops
from conftest import rel_l2
from hallo_b200.spec import UNetConfig
from hallo_b200.synth import host_threads, synth_inputs, synth_refnet_inputs, synth_state_dict, synth_state_dict_2d


@pytest.fixture(scope="module")
def cfg():
    torch.set_num_threads(host_threads())
    return UNetConfig()


@pytest.fixture(scope="module")
def weights(cfg):
    from hallo_b200.engine import PackedWeights
    sd = synth_state_dict(cfg, seed=0)
    return sd, PackedWeights(sd, cfg, torch.device("cpu"), torch.float32)


def _window(inp):
    return dict(encoder_hidden_states=inp["encoder_hidden_states"], audio_embedding=inp["audio_embedding"],
                mask_cond_fea=inp["mask_cond_fea"], full_mask=inp["full_mask"], face_mask=inp["face_mask"],
                lip_mask=inp["lip_mask"], motion_scale=inp["motion_scale"], banks=inp["banks"])


def test_unet3d_plan_matches_oracle(monkeypatch, cfg, weights):
    from hallo_b200.engine import DenoiseEngine
    from oracle import port
    cpu_ops.install(monkeypatch)
    sd, W = weights
    inp = synth_inputs(cfg, 8, 8, 2, seed=3, timestep=500, motion_scale=(1.0, 0.8, 1.2))
    for k in inp["banks"]:
        inp["banks"][k] = inp["banks"][k].float()
    ref = port.unet_forward(sd, cfg, inp)
    eng = DenoiseEngine(W, 8, 8, 2)
    eng.begin_window(**_window(inp))
    eng.set_timestep(inp["timestep"])
    out = eng.forward_only(inp["sample"].float())
    assert rel_l2(out, ref) < 2e-4


def test_denoise_steps_plan_matches_oracle(monkeypatch, cfg, weights):
    """CFG combine + DDIM update + step counter through engine.step() (eager), 2 of 40 steps."""
    from hallo_b200.engine import DenoiseEngine
    from hallo_b200.scheduler import DDIMScheduler
    from oracle import port
    cpu_ops.install(monkeypatch)
    sd, W = weights
    inp = synth_inputs(cfg, 8, 8, 2, seed=4)
    for k in inp["banks"]:
        inp["banks"][k] = inp["banks"][k].float()
    lat0 = inp["sample"][:1].clone()
    ref = port.denoise_loop(sd, cfg, inp, lat0.clone(), 40, 3.5, max_steps=2)
    eng = DenoiseEngine(W, 8, 8, 2)
    eng.begin_window(**_window(inp))
    sch = DDIMScheduler()
    sch.set_timesteps(40)
    eng.set_schedule(sch.timesteps.tolist(), sch.coef_table(), 3.5)
    eng.latents.copy_(lat0)
    eng.step()
    eng.step()
    assert int(eng.step_idx) == 2 and rel_l2(eng.latents, ref) < 2e-4


def test_frame_sharded_plan_matches_unsharded(monkeypatch, cfg, weights):
    """engine._motion_px, the frame <-> pixel swap of a frame-sharded window, with R = 2 ranks simulated as threads:
    every rank runs the product's engine (exchange="nccl" code path) and torch.distributed.all_to_all_single is replaced
    by an in-process exchange (tests/conftest.ThreadGroup).  The gathered sharded output must equal the unsharded
    forward -- frame ownership, pixel slices, motion-frame rows, positional-encoding indices, window slicing."""
    from conftest import ThreadGroup
    from hallo_b200.engine import DenoiseEngine, Shard
    cpu_ops.install(monkeypatch)
    sd, W = weights
    f, R, size = 4, 2, 16
    fl = f // R
    inp = synth_inputs(cfg, size, size, f, seed=31, timestep=600, motion_scale=(1.0, 0.8, 1.2))
    for k in inp["banks"]:
        inp["banks"][k] = inp["banks"][k].float()
    full = DenoiseEngine(W, size, size, f)
    full.begin_window(**_window(inp))
    full.set_timestep(inp["timestep"])
    ref = full.forward_only(inp["sample"].float()).clone()
    grp = ThreadGroup(R, monkeypatch)

    def rank_fn(r):
        frames = tuple(range(r * fl, (r + 1) * fl))
        eng = DenoiseEngine(W, size, size, f, Shard(frames=frames, group=grp, group_size=R, rank_in_group=r, exchange="nccl"))
        if r == 0:       # rank 0 lets the engine slice the full tensors, rank 1 receives caller-sliced ones (local_frames)
            eng.begin_window(**_window(inp))
        else:
            loc = dict(_window(inp))
            loc["audio_embedding"] = inp["audio_embedding"][:, list(frames)]
            loc["mask_cond_fea"] = inp["mask_cond_fea"][:, :, list(frames)]
            lrows = [b * f + g for b in (0, 1) for g in frames]
            for k in ("full_mask", "face_mask", "lip_mask"):
                loc[k] = [t[lrows] for t in inp[k]]
            eng.begin_window(local_frames=True, **loc)
        eng.set_timestep(inp["timestep"])
        return eng.forward_only(inp["sample"][:, :, list(frames)].float()).clone()

    outs = grp.run(rank_fn)
    got = torch.cat(outs, dim=2)
    assert rel_l2(got, ref) < 1e-4


def test_refnet_plan_matches_oracle(monkeypatch, cfg):
    from hallo_b200.refnet import ReferenceNetEngine, ReferenceNetWeights
    from oracle import port
    cpu_ops.install(monkeypatch)
    sd2 = synth_state_dict_2d(cfg)
    W2 = ReferenceNetWeights(sd2, cfg, torch.device("cpu"), torch.float32)
    inp = synth_refnet_inputs(cfg, 8, 8, seed=9)
    ref_out, ref_banks = port.reference_net_forward(sd2, cfg, inp)
    eng = ReferenceNetEngine(W2, 8, 8, inp["sample"].shape[0])
    out, banks = eng.run(inp["sample"], inp["timestep"], inp["encoder_hidden_states"])
    assert rel_l2(out, ref_out) < 2e-4
    assert set(banks) == set(ref_banks) and max(rel_l2(banks[n], ref_banks[n]) for n in banks) < 2e-4
