"""GPU parity ON THE BENCHED CONFIGURATION (BASELINE configs[1]: full-size layout2img UNet at batch 16 on a 6x64x64
latent, MS-VQGAN f8f4 decoder on a 64x64 latent with its N = 4096 attention) against the CPU oracle, plus the
free-running DDIM-200 x 2-stage drift on the full-size UNet.  These are the exact tile / stream-K / attention schedules
`bench.py` times; the small-batch cases of tests/test_gpu_model.py pick different ones.

Tolerance: north star |delta| < 1e-3 fp32 on eps / x_prev / image, VQ indices bit-exact up to near-tie flips of the
argmin (a flip needs two codes within ~1e-5 of each other in distance; counted and bounded)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def l2i(dev):
    from frido_b200 import configs
    model, cfg = configs.build("l2i_coco", dev)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    return model, cfg, sd


def test_unet_eps_at_batch16_64x64_both_stages(dev, l2i):
    """One UNet evaluation per stage at B=16, 64x64 (the bench's plan: BN / stream-K / tile schedules of batch 16), three
    samples of the batch against the oracle evaluated one sample at a time."""
    from oracle import torch_oracle as O
    model, cfg, sd = l2i
    B = 16
    g = torch.Generator().manual_seed(11)
    ctx = torch.randn(B, 26, 640, generator=g)
    for s in (0, 1):
        x = torch.randn(B, 3 * (s + 1), 64, 64, generator=g)
        ts = torch.full((B,), 501, dtype=torch.long)
        e = model.apply_model(x.to(dev), ts.to(dev), ctx.to(dev), stage=s).cpu()
        assert e.shape == (B, 3, 64, 64)
        for i in (0, 7, 15):
            ref = O.unet_forward(sd, x[i:i + 1], ts[i:i + 1], ctx[i:i + 1], s, [3, 3])
            err = (e[i:i + 1] - ref).abs().max().item()
            assert err < 1e-3, (s, i, err)
            assert ref.abs().max().item() > 0.1  # a vacuous (all-zero) output would pass any tolerance


def test_decode_at_64x64_latent_vs_oracle(dev, l2i):
    """Config-2 decode geometry: 64x64 latent -> 256x256 image, mid attention over N = 4096 tokens (d = 512)."""
    from oracle import torch_oracle as O
    model, cfg, sd = l2i
    g = torch.Generator().manual_seed(12)
    z = torch.randn(2, 6, 64, 64, generator=g) * 1.2
    img, codes = model.decode_first_stage(z.to(dev), return_code=True)
    sf = [float(v) for v in model.scale_factor.cpu().tolist()]
    ref, rcodes = O.decode_first_stage(sd, z[:1], [3, 3], sf)
    for a, b in zip(codes, rcodes):
        a = torch.tensor(a)[0].reshape(-1)
        assert (a != b.reshape(-1)).float().mean().item() < 1e-3  # identical arithmetic: flips only on exact near-ties
    err = (img[:1].cpu() - ref).abs().max().item()
    assert err < 1e-3, err
    # fused output formatting (8f.4): the bytes conv_out's epilogue writes == formatting the fp32 image it stores
    import frido_b200 as fb
    for mode in ("np", "pil"):
        u8 = model.decode_first_stage_uint8(z.to(dev), mode=mode)
        assert u8.dtype == torch.uint8 and tuple(u8.shape) == (2, 256, 256, 3)
        assert torch.equal(u8, fb.images_to_uint8(img, mode))
    want = ((ref + 1) * 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)  # custom_to_np on the oracle's image
    got = model.decode_first_stage_uint8(z.to(dev), mode="np")[:1].cpu()
    assert (got.int() - want.int()).abs().max().item() <= 1  # |d image| < 1e-3 is < 0.13 of a grey level: off by one at most


def _drift_case(dev, steps):
    """Free-running DDIM-`steps` x 2 stages on the full-size UNet at a 32x32 latent (BASELINE config 1 geometry), B = 1:
    GPU sampler vs the CPU oracle from the same start noise; no teacher forcing."""
    import frido_b200 as fb
    from frido_b200 import configs
    from oracle import torch_oracle as O
    cfg = configs.get("l2i_coco")
    cfg["model"]["params"]["image_size"] = 32
    cfg["model"]["params"]["unet_config"]["params"]["image_size"] = 32
    torch.manual_seed(0)
    model = fb.FridoDiffusion(**cfg["model"]["params"])
    gw = torch.Generator().manual_seed(1)
    for p in model.parameters():
        if p.numel() > 0 and not p.any():
            p.copy_(torch.randn(p.shape, generator=gw) * 0.02)
    model = model.to(dev).eval()
    model.invalidate_packed_weights()
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items() if k.startswith("model.diffusion_model.")}
    g = torch.Generator().manual_seed(13)
    ctx = torch.randn(1, 26, 640, generator=g)
    x0 = torch.randn(1, 6, 32, 32, generator=g)
    out, _ = fb.DDIMSampler(model).sample(steps, 1, (6, 32, 32), conditioning=ctx.to(dev), num_stage=2, eta=0.0, verbose=False,
                                          init_noise=x0.to(dev))
    ref = O.sample(sd, [3, 3], ctx, x0, steps)
    return out.cpu(), ref


def test_free_running_drift_ddim20(dev):
    out, ref = _drift_case(dev, 20)
    d = (out - ref).abs().max().item()
    assert d < 1e-3, d


@pytest.mark.slow
def test_free_running_drift_ddim200(dev):
    """The full DDIM-200 x 2 run (400 oracle evaluations of the 511 M-parameter UNet on the host: about a minute).
    Prints the drift; `bench.py` reports the same figure as `parity.free_running_absmax` from tools/prof/drift.py."""
    out, ref = _drift_case(dev, 200)
    d = (out - ref).abs().max().item()
    print(f"free-running DDIM-200 x 2 stages, full-size UNet 32x32: final latent |d|max = {d:.3e}, rms = "
          f"{(out - ref).pow(2).mean().sqrt().item():.3e}, latent std = {ref.std().item():.3f}")
    assert d < float(os.environ.get("FRIDO_DRIFT_BOUND", "2e-3")), d
