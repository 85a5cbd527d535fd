"""Model configurations of the shipped Frido YAMLs, restated as Python dicts so
that tests / bench / smoke can build the exact architectures on a box that has
no /root/reference.  Only the `model:` section matters for the sampling path;
values are those of the cited files (checkpoint paths dropped, the out-of-scope
condition encoder replaced by '__is_unconditional__' + a precomputed context).

  l2i_coco : configs/frido/layout2i/frido_f8f4_coco_seg.yaml      (BASELINE config 2)
  t2i_clip : configs/frido/t2i/frido_f16f8_coco_clip.yaml          (BASELINE config 3)
  sg2i_vg  : configs/frido/sg2i/frido_f16f8_vg.yaml                (BASELINE config 4)
  l2i_512  : SURVEY.md §8(d) config 5 (3-scale 512^2, not shipped by the reference)
"""
import copy


def _unet(image_size, split, context_dim):
    c = sum(split)
    return dict(target="frido.modules.diffusionmodules.pyunet.PyUNetModel",
                params=dict(use_split_head=True, split_embed_dim_list=list(split), use_SPADE_norm=True, image_size=image_size,
                            in_channels=c, out_channels=c, model_channels=192, attention_resolutions=[8, 4, 2],
                            num_res_blocks=2, channel_mult=[1, 2, 3, 5], num_head_channels=32, use_spatial_transformer=True,
                            transformer_depth=1, context_dim=context_dim, num_stage=len(split)))


def _first_stage(embed_dim, n_embed, ed_ch_mult, dd_ch_mult, attn_res, resolution=256):
    return dict(target="taming.models.msvqgan.VQModelInterface",
                params=dict(ckpt_path=None, embed_dim=list(embed_dim), n_embed=list(n_embed), init_normal=True,
                            edconfig=dict(multiscale=len(embed_dim), double_z=False, z_channels=list(embed_dim),
                                          resolution=resolution, in_channels=3, out_ch=3, ch=128, ch_mult=list(ed_ch_mult),
                                          num_res_blocks=2, attn_resolutions=[attn_res], dropout=0.0),
                            ddconfig=dict(double_z=False, z_channels=sum(embed_dim), resolution=resolution, in_channels=3,
                                          out_ch=3, ch=128, ch_mult=list(dd_ch_mult), num_res_blocks=2,
                                          attn_resolutions=[attn_res], dropout=0.0),
                            lossconfig=dict(target="taming.modules.losses.DummyLoss")))


def _model(image_size, split, context_dim, first_stage, cond_stage_key):
    return dict(target="frido.models.diffusion.frido.FridoDiffusion",
                params=dict(adopted_scale_factor=True, noise_mix_ratio=0.1, first_stage_key="image",
                            cond_stage_key=cond_stage_key, linear_start=0.0015, linear_end=0.0155, num_timesteps_cond=1,
                            log_every_t=200, timesteps=1000, loss_type="l1", image_size=image_size, channels=sum(split),
                            cond_stage_trainable=False, conditioning_key="crossattn", scale_by_std=True, monitor="val/loss",
                            use_ema=False, stage_loss_ratio=[1.0 / len(split)] * len(split),
                            unet_config=_unet(image_size, split, context_dim), first_stage_config=first_stage,
                            cond_stage_config="__is_unconditional__", plot_sample=False, plot_inpaint=False,
                            plot_denoise_rows=False, plot_progressive_rows=False, plot_diffusion_rows=False,
                            plot_quantize_denoised=True))


CONFIGS = {
    # latent 6x64x64, ctx [B,26,640] (8 objects x 3 tokens + 2 crop tokens through the 640-d BERTEmbedder)
    "l2i_coco": dict(model=_model(64, [3, 3], 640, _first_stage([3, 3], [4096, 4096], [1, 1, 2, 4], [1, 2, 4], 64), "objects_bbox"),
                     latent=(6, 64, 64), ctx=(26, 640), sampler="ddim", steps=200, cfg_scale=1.0, batch=16),
    # latent 8x32x32, ctx [B,1,768] (CLIP pooled text)
    "t2i_clip": dict(model=_model(32, [4, 4], 768, _first_stage([4, 4], [8192, 8192], [1, 1, 2, 2, 4], [1, 1, 2, 4], 32), "caption"),
                     latent=(8, 32, 32), ctx=(1, 768), sampler="plms", steps=100, cfg_scale=1.0, batch=32),
    # latent 8x32x32, ctx [B,180,640] (scene-graph tokens padded to max_seq_len 180)
    "sg2i_vg": dict(model=_model(32, [4, 4], 640, _first_stage([4, 4], [8192, 8192], [1, 1, 2, 2, 4], [1, 1, 2, 4], 32), "caption"),
                    latent=(8, 32, 32), ctx=(180, 640), sampler="ddim", steps=200, cfg_scale=1.0, batch=8),
    # 3-scale 512^2 extrapolation: latent 9x128x128, ctx [B,92,640]
    "l2i_512": dict(model=_model(128, [3, 3, 3], 640,
                                 _first_stage([3, 3, 3], [4096, 4096, 4096], [1, 1, 2, 2, 4], [1, 2, 4], 128, resolution=512),
                                 "objects_bbox"),
                    latent=(9, 128, 128), ctx=(92, 640), sampler="ddim", steps=250, cfg_scale=1.0, batch=16),
}


def get(name):
    return copy.deepcopy(CONFIGS[name])


def build(name, device, seed=0, scale_factor=(0.8, 1.3, 1.1)):
    """Instantiate a config with SYNTHETIC weights (no checkpoints exist here):
    default init under torch.manual_seed(seed), every all-zero tensor (`zero_module`: ResBlock conv2,
    transformer proj_out, out heads) re-drawn from N(0, 0.02^2) — otherwise the UNet outputs exactly 0 —
    codebooks N(0,1), per-scale scale_factor set to exercise the decode rescale."""
    import torch

    from .diffusion import FridoDiffusion

    cfg = get(name)
    torch.manual_seed(seed)
    model = FridoDiffusion(**cfg["model"]["params"])
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in model.parameters():
            if p.numel() > 0 and not p.any():
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        n = len(model.first_stage_model.embed_dim)
        model.scale_factor.copy_(torch.tensor(list(scale_factor[:n])))
    model = model.to(device).eval()
    model.invalidate_packed_weights()
    return model, cfg
