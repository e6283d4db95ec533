"""ORACLE (test infrastructure only): the reference's Dreambooth training step in plain PyTorch.

Restates training_scripts/train_lora_dreambooth.py:651-676 (AdamW param groups) and :811-888
(the step) with the stock torch pieces the reference itself calls -- torch.optim.AdamW,
torch.nn.utils.clip_grad_norm_, F.mse_loss -- over RefLoraSite modules (oracle/ref_modules.py).
The host UNet / text encoder / noiser objects are passed in by the caller (tests, bench.py);
this file imports nothing from the product package.

Pinned: `forward_loss` reproduces the losses of the reference's own `loss_step`
(lora_diffusion/cli_lora_pti.py:260-370; plain, t_mutliplier, masked, inpainting) to 2e-6 relative
-- tests/golden/pti_loss_step.pt, written by scripts/make_golden.py::gen_loss_step, checked in
tests/test_oracle_golden.py; `step` reproduces 3 iterations of the reference's real
`perform_tuning` loop (cli_lora_pti.py:545-680: per-step loss to 5e-6, every LoRA factor afterwards
to 2e-7 -- tests/golden/pti_perform_tuning.pt, gen_tuning); the clip + AdamW trajectory alone is
also pinned by tests/golden/adamw_clip.pt.
"""
import itertools
from typing import List, Optional

import torch
import torch.nn.functional as F


class RefDreamboothStep:
    def __init__(self, unet, text_encoder, noiser, unet_sites: List, text_sites: Optional[List],
                 lr=1e-4, lr_text=5e-5, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8,
                 max_grad_norm=1.0, t_multiplier: float = 1.0, autocast_dtype=None, capturable: bool = False):
        self.unet, self.text_encoder, self.noiser = unet, text_encoder, noiser
        self.max_grad_norm = max_grad_norm
        self.t_multiplier = t_multiplier
        self.autocast_dtype = autocast_dtype
        self.train_text = bool(text_sites)

        def factors(sites):  # reference order per site: up then down (lora.py:298-299)
            return list(itertools.chain(*[(s.up, s.down) for s in sites]))

        for p in itertools.chain(unet.parameters(), text_encoder.parameters()):
            p.requires_grad_(False)
        self.unet_params = factors(unet_sites)
        self.text_params = factors(text_sites) if text_sites else []
        for p in self.unet_params + self.text_params:
            p.requires_grad_(True)
        groups = [{"params": self.unet_params, "lr": lr}]
        if self.text_params:
            groups.append({"params": self.text_params, "lr": lr_text})
        # capturable=True only so that bench.py can ALSO time this step as a CUDA-graph replay (the
        # reference itself runs eager; the graphed figure is the generous comparator)
        self.opt = torch.optim.AdamW(groups, lr=lr, betas=betas, weight_decay=weight_decay, eps=eps,
                                     **({"capturable": True} if capturable else {}))
        unet.train()
        text_encoder.train()

    def forward_loss(self, latents, input_ids, noise, timesteps, loss_mask=None, mask_temperature=1.0,
                     inpaint=None, prior_loss_weight=None):
        """loss of cli_lora_pti.py:260-370 (`loss_step`, epsilon prediction, cached latents).
        loss_mask: [B,1,8h,8w] image-resolution mask (`batch["mask"]`, :340-368) or None;
        inpaint: (mask [B,1,h,w], masked_image_latents [B,4,h,w]) for the 9-channel UNet (:279-313);
        prior_loss_weight: Dreambooth prior preservation (train_lora_dreambooth.py:855-873) or None."""
        noisy = self.noiser.add_noise(latents, noise, timesteps)
        model_in = noisy if inpaint is None else torch.cat([noisy, inpaint[0], inpaint[1]], dim=1)
        dev = latents.device.type
        ctx = torch.autocast(dev, dtype=self.autocast_dtype) if self.autocast_dtype else torch.autocast(dev, enabled=False)
        with ctx:
            if self.train_text:
                ehs = self.text_encoder(input_ids)[0]
            else:
                with torch.no_grad():
                    ehs = self.text_encoder(input_ids)[0]
            mdt = next(self.unet.parameters()).dtype
            pred = self.unet(model_in.to(mdt), timesteps, ehs.to(mdt)).sample
        target = noise
        if loss_mask is not None:
            m = loss_mask.to(pred.device).reshape(pred.shape[0], 1, pred.shape[2] * 8, pred.shape[3] * 8)
            m = F.interpolate(m.float(), size=pred.shape[-2:], mode="nearest")
            m = (m + 0.01).pow(mask_temperature)
            m = m / m.max()
            pred, target = pred * m, target * m
        if prior_loss_weight is not None:
            # train_lora_dreambooth.py:855-873: batch = [instance ; class], two losses
            pred, pred_prior = torch.chunk(pred, 2, dim=0)
            target, target_prior = torch.chunk(target, 2, dim=0)
            inst = F.mse_loss(pred.float(), target.float(), reduction="none").mean([1, 2, 3]).mean()
            return inst + prior_loss_weight * F.mse_loss(pred_prior.float(), target_prior.float(), reduction="mean")
        if loss_mask is not None:
            return F.mse_loss(pred.float(), target.float(), reduction="none").mean([1, 2, 3]).mean()
        return F.mse_loss(pred.float(), target.float(), reduction="mean")

    def step(self, latents, input_ids, noise=None, timesteps=None, **loss_kw):
        """noise/timesteps may be supplied (parity tests); otherwise drawn like the reference
        (noise first, then timesteps: cli_lora_pti.py:295-304, train_lora_dreambooth.py:822-833)."""
        if noise is None:
            noise = torch.randn_like(latents)
        if timesteps is None:
            t_max = int(self.noiser.num_train_timesteps * self.t_multiplier)
            timesteps = torch.randint(0, t_max, (latents.shape[0],), device=latents.device).long()
        loss = self.forward_loss(latents, input_ids, noise, timesteps, **loss_kw)
        loss.backward()
        if self.max_grad_norm:
            torch.nn.utils.clip_grad_norm_(self.unet_params + self.text_params, self.max_grad_norm)
        self.opt.step()
        self.opt.zero_grad()
        return loss.detach()
