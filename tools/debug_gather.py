"""2-rank debug of the frame-sharded temporal K/V gather: both ranks take CFG half 1 and one half of the frames
(not a valid denoising step -- only exercises DenoiseEngine._forward with group_size 2), eagerly and graph-captured.
    torchrun --nproc-per-node 2 tools/debug_gather.py"""
import faulthandler
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(int(os.environ.get('WATCHDOG', 120)), repeat=False, exit=True)

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from hallo_b200.engine import DenoiseEngine, PackedWeights, Shard
from hallo_b200.spec import UNetConfig
from hallo_b200.synth import synth_inputs, synth_state_dict, host_threads

torch.set_num_threads(max(1, host_threads() // world))
cfg = UNetConfig()
t0 = time.time()
from hallo_b200.spec import param_spec, sinusoid_pe
torch.manual_seed(0)          # same seed on every rank -> identical weights, generated on the GPU in a second
sd = {}
for key, shape, kind in param_spec(cfg):
    if kind == "pe":
        sd[key] = sinusoid_pe(shape[1], shape[2]).to(dev)
    elif kind in ("w", "zero_w"):
        fan = 1
        for q in shape[1:]:
            fan *= q
        sd[key] = torch.randn(shape, device=dev) * fan ** -0.5
    elif kind == "norm_w":
        sd[key] = 1 + 0.1 * torch.randn(shape, device=dev)
    else:
        sd[key] = 0.05 * torch.randn(shape, device=dev)
W = PackedWeights(sd, cfg, dev, torch.float16)
del sd
print(rank, "weights", round(time.time() - t0, 1), flush=True)
size, f = int(os.environ.get("SIZE", 32)), 16
inp = synth_inputs(cfg, size, size, f, seed=42)
grp = dist.new_group(list(range(world)))
fl = f // world
sh = Shard(halves=(1,), frames=tuple(range(rank * fl, (rank + 1) * fl)), group=grp, group_size=world,
           world=dist.group.WORLD, world_size=world, rank_in_group=rank)
eng = DenoiseEngine(W, size, size, f, sh)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items() if k not in ("banks", "sample", "timestep", "motion_scale", "full_mask", "face_mask", "lip_mask")}
eng.begin_window(encoder_hidden_states=inp["encoder_hidden_states"].to(dev), audio_embedding=inp["audio_embedding"].to(dev),
                 mask_cond_fea=inp["mask_cond_fea"].to(dev), full_mask=[m.to(dev) for m in inp["full_mask"]],
                 face_mask=[m.to(dev) for m in inp["face_mask"]], lip_mask=[m.to(dev) for m in inp["lip_mask"]],
                 motion_scale=inp["motion_scale"], banks={k: v.to(dev) for k, v in inp["banks"].items()})
eng.set_timestep(999.0)
eng.latents.copy_(inp["sample"][:1, :, list(sh.frames)].to(dev))
print(rank, "begin_window done", flush=True)
for i in range(2):
    eng._forward()
    torch.cuda.synchronize()
    print(rank, "eager forward", i, "ok", float(eng.model_out.float().abs().mean()), flush=True)
dist.barrier()
# single-rank reference of the same rows: full frames on this rank, compare the local frames
eng1 = DenoiseEngine(W, size, size, f, Shard(halves=(1,), frames=tuple(range(f))))
eng1.begin_window(encoder_hidden_states=inp["encoder_hidden_states"].to(dev), audio_embedding=inp["audio_embedding"].to(dev),
                  mask_cond_fea=inp["mask_cond_fea"].to(dev), full_mask=[m.to(dev) for m in inp["full_mask"]],
                  face_mask=[m.to(dev) for m in inp["face_mask"]], lip_mask=[m.to(dev) for m in inp["lip_mask"]],
                  motion_scale=inp["motion_scale"], banks={k: v.to(dev) for k, v in inp["banks"].items()})
eng1.set_timestep(999.0)
eng1.latents.copy_(inp["sample"][:1].to(dev))
eng1._forward()
torch.cuda.synchronize()
L0 = size * size
ref = eng1.model_out.view(f, L0, 8)[rank * fl:(rank + 1) * fl].float()
got = eng.model_out.view(fl, L0, 8).float()
print(rank, "sharded vs unsharded rel err", float((got - ref).norm() / ref.norm()), flush=True)
dist.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, capture_error_mode="thread_local"):
    eng._forward()
print(rank, "captured", flush=True)
for i in range(3):
    g.replay()
torch.cuda.synchronize()
print(rank, "graph replay ok", float(eng.model_out.float().abs().mean()), flush=True)
dist.barrier()
sys.stdout.flush()
os._exit(0)   # destroy_process_group() hangs while graphs holding NCCL collectives are alive
