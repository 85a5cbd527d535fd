"""DDIMSampler / PLMSSampler — drop-in mirrors of frido/models/diffusion/ddim.py
and plms.py (same constructor, `make_schedule`, `sample(...)` signature and
return value), re-designed around device-resident sampler state:

  * one CUDA graph per stage = [timestep broadcast] + [UNet step program] +
    [fused x_{t-1} update]; the step counter, timestep table and coefficient
    table live on the device, so the T steps are T graph replays with no host
    sync (the reference does 3 implicit D2H syncs per step, ddim.py:237-240);
  * the UNet plan's input buffer IS the sampler's x: the update kernel writes
    x_{t-1} in place;
  * classifier-free guidance runs cond and uncond as one 2B batch and the
    combine (ddim.py:226) is fused into the update kernel;
  * PLMS keeps its eps history in a device ring buffer.

Quirks kept for drop-in fidelity: passing `x_T` skips stage 0 entirely
(ddim.py:150-152).  Use `init_noise=` to inject the start noise instead.
"""
import numpy as np
import torch

from . import _lib as L
from .program import Program


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """util.py:46-60."""
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * 0.8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    steps_out = ddim_timesteps + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """util.py:63-74 with the reference's mixed fp32/fp64 evaluation order:
    `alphas` fp32, `alphas_prev` fp64 values of fp32 numbers, `1 - alphas` in fp32 and
    `ndarray / tensor` = other * tensor.reciprocal() (fp32 reciprocal), rest fp64."""
    acp = np.asarray(alphacums, dtype=np.float32)
    alphas = acp[ddim_timesteps]
    alphas_prev = np.asarray([acp[0]] + acp[ddim_timesteps[:-1]].tolist(), dtype=np.float64)
    recip = (np.float32(1.0) / (np.float32(1.0) - alphas)).astype(np.float64)
    sigmas = eta * np.sqrt((1 - alphas_prev) * recip * (1 - alphas.astype(np.float64) / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, this results in the following sigma_t schedule {sigmas}")
    return sigmas, alphas, alphas_prev


class _SamplerBase(object):
    KIND = "ddim"

    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self._stage_cache = {}

    def register_buffer(self, name, attr):
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        if self.KIND == "plms" and ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")  # plms.py:25-26
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose)
        acp = self.model.alphas_cumprod.detach().float().cpu().numpy()
        assert acp.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        sig, a, ap = make_ddim_sampling_parameters(acp, self.ddim_timesteps, ddim_eta, verbose)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sig, a, ap
        self.ddim_sqrt_one_minus_alphas = np.sqrt(np.float32(1.0) - a)  # fp32 (ddim.py:50)
        self.ddim_eta = ddim_eta
        T = len(self.ddim_timesteps)
        # device tables in step order i (index = T-1-i): {a_t, a_prev, sigma, sqrt(1-a_t)} as fp32
        tab = np.stack([a.astype(np.float32), ap.astype(np.float32), sig.astype(np.float32),
                        self.ddim_sqrt_one_minus_alphas.astype(np.float32)], 1)[::-1].copy()
        dev = self.model.device
        self._coef = torch.from_numpy(tab).to(dev)
        self._t_table = torch.from_numpy(np.flip(self.ddim_timesteps).astype(np.int64).copy()).to(dev)
        self._T = T

    # ------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, num_stage=1, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0.0, mask=None, x0=None, temperature=1.0, noise_dropout=0.0,
               score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100,
               unconditional_guidance_scale=1.0, unconditional_conditioning=None, init_noise=None, noise_sequence=None,
               seed=None, mask_noise_sequence=None, **kwargs):
        if conditioning is not None:
            cbs = (conditioning[list(conditioning.keys())[0]] if isinstance(conditioning, dict) else conditioning).shape[0]
            if cbs != batch_size:
                print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
        if quantize_x0 or score_corrector is not None or noise_dropout > 0.0:
            raise NotImplementedError("quantize_x0, score_corrector and noise_dropout are outside the B200 sampling hot "
                                      "path (SURVEY.md §2)")
        if mask is not None:
            assert x0 is not None  # ddim.py:159
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        if verbose:
            print(f"Data shape for {self.KIND.upper()} sampling is {(batch_size, C, H, W)}, eta {eta}")
        return self._sampling(conditioning, (batch_size, C, H, W), num_stage, x_T, log_every_t, temperature,
                              unconditional_guidance_scale, unconditional_conditioning, init_noise, noise_sequence, seed,
                              callback, img_callback, mask, x0, mask_noise_sequence)

    def _cond_tensor(self, cond):
        if isinstance(cond, dict):
            c = cond.get("c_crossattn")
            cond = torch.cat(c, 1) if isinstance(c, (list, tuple)) else c
        elif isinstance(cond, (list, tuple)):
            cond = torch.cat(list(cond), 1)
        return cond

    def _sampling(self, cond, shape, num_stage, x_T, log_every_t, temperature, cfg_scale, uc, init_noise, noise_sequence,
                  seed, callback, img_callback, mask=None, x0=None, mask_noise_sequence=None):
        model = self.model
        dev = model.betas.device
        if dev.type != "cuda":
            raise L.FridoError("sampling runs on a CUDA device only (no CPU path)")
        unet = model.model.diffusion_model
        unet.invalidate_if_changed()  # EMA swap may have rewritten the weights in place (sample_diffusion.py:187)
        B, C, H, W = shape
        split = list(model.split_embed_dim_list) if getattr(model, "use_split_head", False) else [C]
        if not getattr(model, "use_split_head", False):
            raise NotImplementedError("only split-head Frido models are supported")
        if x_T is not None:
            img = x_T.clone().to(dev, torch.float32)
        elif init_noise is not None:
            img = init_noise.clone().to(dev, torch.float32)
        else:
            img = torch.randn(shape, device=dev)
        cond = self._cond_tensor(cond).to(dev, torch.float32)
        use_cfg = cfg_scale != 1.0
        if use_cfg:
            assert uc is not None
            uc = self._cond_tensor(uc).to(dev, torch.float32)
        T = self._T
        start = img.clone()  # entry 0 stays x_T like the reference's (it rebinds img every step; we update in place)
        intermediates = {"x_inter": [start], "pred_x0": [start]}
        self.num_stage = num_stage
        # host-side draw from torch's CPU generator (no device sync): the Philox seed of this call
        seed = int(torch.randint(0, 2**62, (1,), device="cpu").item()) if seed is None else int(seed)
        nk = mk = 0
        self.launches = 0
        for s in range(num_stage):
            c_start, c_end = sum(split[:s]), sum(split[: s + 1])
            if x_T is not None and s == 0:
                print("Find x_T is not None. Auto adopt x_T into stage 0.")  # ddim.py:150-152
                continue
            masked = mask is not None
            if masked:
                # ddim.py:160-161: `img_orig * mask + (1. - mask) * img` with img = the channels of this stage; the reference
                # leaves shape errors to broadcasting (with split heads only the stage that carries all of x0's channels works)
                if tuple(torch.broadcast_shapes(tuple(x0.shape), tuple(mask.shape), (B, c_end, H, W))) != (B, c_end, H, W):
                    raise RuntimeError(f"The size of tensor a ({x0.shape[1]}) must match the size of tensor b ({c_end}) at "
                                       f"non-singleton dimension 1 (mask/x0 blend at stage {s}, ddim.py:161)")
            st = self._stage(unet, s, B, H, W, cond.shape[1], use_cfg, cfg_scale, c_start, c_end, temperature, masked)
            plan = st["plan"]
            plan.repack_if_stale()
            x = plan.x_in
            if masked:
                st["x0"].copy_(x0.to(dev, torch.float32).expand(B, c_end, H, W))
                st["mask"].copy_(mask.to(dev, torch.float32).expand(B, c_end, H, W))
            x[:B].copy_(img[:, :c_end])
            plan.ctx[:B].copy_(cond)
            if use_cfg:
                x[B:].copy_(img[:, :c_end])
                plan.ctx[B:].copy_(uc)
            st["step"].zero_()
            st["seed_dev"].copy_(torch.tensor([(seed + 7919 * s) & (2**63 - 1)], dtype=torch.int64), non_blocking=True)
            plan.prologue.run()
            self.launches += len(plan.prologue)
            first = 0
            for i in range(T):
                index = T - i - 1
                inj = None
                if noise_sequence is not None:
                    inj = noise_sequence[nk]
                    nk += 1
                if inj is not None:
                    st["noise"].copy_(inj[:, :c_end])
                minj = None
                if masked and mask_noise_sequence is not None:
                    minj = mask_noise_sequence[mk]
                    mk += 1
                    st["mnoise"].copy_(minj[:, :c_end])
                    if inj is None and self.ddim_eta != 0:
                        st["noise"].normal_()  # the injected-blend program reads the step noise from this buffer too
                if self.KIND == "plms" and i == 0:
                    if masked:  # the blend (and the per-step SPADE maps it invalidates) precede both half steps (plms.py:162-165)
                        pre = st["blend_inj"] if minj is not None else st["blend"]
                        pre.run()
                        self.launches += len(pre)
                    st["x_orig"].copy_(x[:B])
                    st["first_a"].run()
                    st["first_b"].run()
                    self.launches += len(st["first_a"]) + len(st["first_b"])
                else:
                    if masked:
                        prog = st["mask_inj"] if (inj is not None or minj is not None) else st["mask_main"]
                    else:
                        prog = st["inj"] if inj is not None else st["main"]
                    prog.replay()
                    self.launches += len(prog)
                if callback:
                    callback(i)
                if img_callback:
                    img_callback(st["pred_x0"].clone(), i)
                if index % log_every_t == 0 or index == T - 1:
                    intermediates["x_inter"].append(x[:B].clone())
                    intermediates["pred_x0"].append(st["pred_x0"].clone())
            img[:, :c_end].copy_(x[:B])
            if num_stage != 1:
                n = num_stage - s - 1
                if n > 0:
                    snap = Program(dev, "snap")
                    snap.snap(img, B=B, Ctot=C, H=H, W=W, c_start=c_start, c_end=c_end, n=n)
                    snap.run()
                    self.launches += 1
        out = img if num_stage == len(split) else img[:, : sum(split[:num_stage])]
        # the reference self-synchronises three times a step (ddim.py:237-240), so callers time sample() with the host clock
        # (sample_diffusion.py:188-205): hand the result back only when the device has produced it
        torch.cuda.current_stream(dev).synchronize()
        return out, intermediates

    def _stage(self, unet, s, B, H, W, Lc, use_cfg, cfg_scale, c_start, c_end, temperature, masked=False):
        """Build (once) the per-stage programs and device state."""
        key = (s, B, H, W, Lc, use_cfg, float(cfg_scale), self._T, float(temperature), self.KIND, masked)
        st = self._stage_cache.get(key)
        if st is not None:
            st["coef"].copy_(self._coef)
            st["t_table"].copy_(self._t_table)
            return st
        dev = self.model.betas.device
        Bn = 2 * B if use_cfg else B
        plan = unet.plan(s, Bn, H, W, Lc)
        HW = H * W
        c_act = c_end - c_start
        f32 = dict(dtype=torch.float32, device=dev)
        st = dict(plan=plan, step=torch.zeros(1, dtype=torch.int32, device=dev), coef=self._coef.clone(),
                  t_table=self._t_table.clone(), pred_x0=torch.zeros(B, c_end, H, W, **f32),
                  noise=torch.zeros(B, c_end, H, W, **f32), seed_dev=torch.zeros(1, dtype=torch.int64, device=dev))
        x = plan.x_in
        eps_c = plan.eps[:B]
        eps_u = plan.eps[B:] if use_cfg else None
        x_dup = x[B:] if use_cfg else None
        plms = self.KIND == "plms"
        if plms:
            st["hist"] = torch.zeros(3, B, c_act, H, W, **f32)
            st["eps_save"] = torch.zeros(B, c_act, H, W, **f32)
            st["x_orig"] = torch.zeros(B, c_end, H, W, **f32)
        common = dict(B=B, c_start=c_start, c_end=c_end, HW=HW, eps_uncond=eps_u, cfg_scale=float(cfg_scale),
                      temperature=float(temperature), x_dup=x_dup, pred_x0=st["pred_x0"],
                      hist=st.get("hist"), eps_save=st.get("eps_save"))

        def full(name, use_next, upd_kwargs, front=None):
            pre = Program(dev, name)
            pre.step_begin(st["step"], st["t_table"], plan.ts, B=Bn, T=self._T, use_next=use_next)
            post = Program(dev, name)
            post.update(eps=eps_c, coef=st["coef"], step=st["step"], **upd_kwargs, **common)
            prog = Program(dev, name)
            f_ops, f_tags = (front.ops, front.tags) if front is not None else ([], [])
            prog.ops = f_ops + pre.ops + plan.step.ops + post.ops
            prog.tags = f_tags + pre.tags + plan.step.tags + post.tags
            prog.keep = [pre, post, plan, front]
            prog.flops = plan.step.flops + (front.flops if front is not None else 0)
            return prog

        def blend_front(name, noise):
            """mask / x0 mode: the blend rewrites ALL channels before every step (ddim.py:158-161), so the coarse groups are no
            longer frozen and the step-invariant hoist of the SPADE maps does not hold: the plan's prologue (h_cond, the
            gamma|beta maps of every norm site) re-runs in front of each step."""
            fr = Program(dev, name)
            fr.blend(x[:B], st["x0"], st["mask"], self.model.sqrt_alphas_cumprod, self.model.sqrt_one_minus_alphas_cumprod,
                     st["step"], st["t_table"], B=B, Cdim=c_end, HW=HW, T=self._T, noise=noise, x_dup=x_dup, seed=0xB1E4D + s,
                     seed_dev=st["seed_dev"])
            if plan.c_cond:
                fr.ops = fr.ops + plan.prologue.ops
                fr.tags = fr.tags + plan.prologue.tags
                fr.flops += plan.prologue.flops
                fr.keep.append(plan.prologue)
            return fr

        order = 4 if plms else 0
        st["main"] = full(f"{self.KIND}.s{s}", 0, dict(x=x[:B], x_prev=x[:B], plms_order=order, plms_mode=0, advance=1,
                                                         noise=None, seed=0x5EED + s, seed_dev=st["seed_dev"]))
        st["inj"] = full(f"{self.KIND}.s{s}.inj", 0, dict(x=x[:B], x_prev=x[:B], plms_order=order, plms_mode=0, advance=1,
                                                           noise=st["noise"]))
        if plms:
            st["first_a"] = full("plms.first_a", 0, dict(x=st["x_orig"], x_prev=x[:B], plms_order=4, plms_mode=1, advance=0))
            st["first_b"] = full("plms.first_b", 1, dict(x=st["x_orig"], x_prev=x[:B], plms_order=4, plms_mode=2, advance=1))
        if masked:
            st["x0"] = torch.zeros(B, c_end, H, W, **f32)
            st["mask"] = torch.zeros(B, c_end, H, W, **f32)
            st["mnoise"] = torch.zeros(B, c_end, H, W, **f32)
            st["blend"] = blend_front(f"{self.KIND}.s{s}.blend", None)
            st["blend_inj"] = blend_front(f"{self.KIND}.s{s}.blend.inj", st["mnoise"])
            st["mask_main"] = full(f"{self.KIND}.s{s}.mask", 0, dict(x=x[:B], x_prev=x[:B], plms_order=order, plms_mode=0, advance=1,
                                                                      noise=None, seed=0x5EED + s, seed_dev=st["seed_dev"]),
                                   front=st["blend"])
            st["mask_inj"] = full(f"{self.KIND}.s{s}.mask.inj", 0, dict(x=x[:B], x_prev=x[:B], plms_order=order, plms_mode=0,
                                                                         advance=1, noise=st["noise"]), front=st["blend_inj"])
            st["mask_main"].capture()
        st["main"].capture()
        self._stage_cache[key] = st
        return st


class DDIMSampler(_SamplerBase):
    KIND = "ddim"

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, s, index, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
                      temperature=1.0, noise=None, **kwargs):
        """Single step with the reference's signature (ddim.py:189-273): returns (x_prev, pred_x0)."""
        model = self.model
        dev = x.device
        B, _, H, W = x.shape
        split = list(model.split_embed_dim_list)
        c_start, c_end = sum(split[:s]), sum(split[: s + 1])
        e_t = model.apply_model(x, t, c, stage=s)
        e_u = None
        if unconditional_guidance_scale != 1.0:
            e_u = model.apply_model(x, t, unconditional_conditioning, stage=s)
        T = self._T
        step = torch.tensor([T - 1 - index], dtype=torch.int32, device=dev)
        xp = torch.empty_like(x[:, :c_end])
        p0 = torch.empty_like(xp)
        prog = Program(dev, "p_sample_ddim")
        prog.update(x[:, :c_end].contiguous(), e_t.contiguous(), self._coef, step, xp, B=B, c_start=c_start, c_end=c_end,
                    HW=H * W, eps_uncond=None if e_u is None else e_u.contiguous(), cfg_scale=float(unconditional_guidance_scale),
                    advance=0, noise=noise, temperature=float(temperature), pred_x0=p0)
        prog.run()
        return xp, p0


class PLMSSampler(_SamplerBase):
    KIND = "plms"
