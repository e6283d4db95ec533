"""N-GPU data-parallel parity (run under torchrun): the LoRA parameters after K steps on N ranks
(each rank its own sample, one all-reduce per step, 1/N folded into the fused AdamW) equal, to
fp32 reduction-order noise, those of ONE process that accumulates the same N samples' gradients
and applies the same fused step with inv_world = 1/N. Tiny host models, dropout 0, eager steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

import lora_b200 as L
from lora_b200.host.clip import build_text_encoder
from lora_b200.host.unet_sd15 import UNet2DConditionModel, UNetConfig
from lora_b200.train import LoraTrainStep, StepConfig


def build(dev):
    torch.manual_seed(0)
    unet = UNet2DConditionModel(UNetConfig.tiny()).to(dev).to(torch.bfloat16)
    text = build_text_encoder(tiny=True).to(dev).to(torch.bfloat16)
    unet.requires_grad_(False); text.requires_grad_(False)
    L.inject_trainable_lora(unet, r=4)
    L.inject_trainable_lora(text, target_replace_module={"CLIPAttention"}, r=4)
    g = torch.Generator(device=dev).manual_seed(1)
    for m in list(unet.modules()) + list(text.modules()):
        if type(m).__name__ == "LoraInjectedLinear":
            m.lora_up.weight.data.normal_(0, 0.05, generator=g)
    return unet, text


def sample(rank, dev):
    g = torch.Generator().manual_seed(1000 + rank)
    return (torch.randn(1, 4, 16, 16, generator=g).to(dev) * 0.18215,
            torch.randint(0, 1000, (1, 77), generator=g).to(dev),
            torch.randn(1, 4, 16, 16, generator=g).to(dev),
            torch.randint(0, 1000, (1,), generator=g).to(dev))


def fwd_bwd(tr, unet, text, lat, ids, noise, t):
    noisy = tr.noiser.add_noise(lat, noise, t).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    pred = unet(noisy, t, text(ids)[0]).sample
    loss = F.mse_loss(pred.float(), noise.float())
    loss.backward()
    return float(loss)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    cfg = StepConfig(use_cuda_graph=False, peer_allreduce=os.environ.get("LB_DP_NCCL", "0") != "1")
    # (a) distributed: each rank one sample
    unet, text = build(dev)
    tr = LoraTrainStep(unet, text, cfg, latent_shape=(1, 4, 16, 16), device=dev)
    for step in range(3):
        fwd_bwd(tr, unet, text, *sample(rank + 10 * step, dev))
        ws = tr.arena.allreduce_grads()
        assert ws == world
        tr.arena.step(world_size=world)
    p_dist = tr.arena.p.clone()
    # (b) single process emulation on every rank: accumulate all N samples, inv_world = 1/N
    unet2, text2 = build(dev)
    cfg2 = StepConfig(use_cuda_graph=False, peer_allreduce=False)     # local emulation: no exchange at all
    tr2 = LoraTrainStep(unet2, text2, cfg2, latent_shape=(1, 4, 16, 16), device=dev)
    tr2._world = 1
    for step in range(3):
        for r in range(world):
            fwd_bwd(tr2, unet2, text2, *sample(r + 10 * step, dev))
        tr2.arena.step(world_size=world)
    torch.cuda.synchronize()
    err = float((p_dist - tr2.arena.p).norm() / tr2.arena.p.norm())
    # replicas identical across ranks
    ref = p_dist.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, p_dist))
    flags = torch.tensor([err, 0.0 if same else 1.0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"dp_parity world={world} exchange={'nvlink-peer (lb_optim_step_dp)' if tr.peer_allreduce else 'nccl'} "
              f"rel_err_vs_single_process={float(flags[0]):.3e} replicas_identical={float(flags[1]) == 0.0}")
        assert float(flags[0]) < 5e-3 and float(flags[1]) == 0.0
        print("dp_parity OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
