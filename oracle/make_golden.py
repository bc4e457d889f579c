"""Generate tests/golden/*.npz by executing the UNMODIFIED reference modules (TEST INFRASTRUCTURE ONLY).

Run in the build container (needs /root/reference):  python oracle/make_golden.py
The reference draws noise from the global RNG inside p_sample / SineGen; here those draws are replaced by the
seeded tensors of oracle/synth.py (monkey-patching ``noise_like`` and the ``torch`` name inside models/source.py),
so the fixtures are a pure function of (seed, shapes).  Weights come from oracle/synth.py (same seeds on every
machine), so only inputs' seeds and the reference OUTPUTS are stored.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

DIFF_CASES = [dict(seed=7, B=2, T=40, K=100), dict(seed=8, B=1, T=150, K=100)]
EPS_CASES = [dict(seed=21, B=2, T=50, t=99), dict(seed=22, B=2, T=50, t=50), dict(seed=23, B=1, T=133, t=0)]
VOC_CASES = [dict(seed=11, B=2, T=16), dict(seed=12, B=1, T=40)]
K_STEP, MAX_BETA = 100, 0.06


def build_reference_sampler(ns):
    sd = synth.diffnet_state(1234)
    net = ns.DiffNet(80)
    net.load_state_dict(sd, strict=True)
    net.eval()
    gd = ns.gd
    gd.FastSpeech2 = lambda *a, **k: torch.nn.Identity()   # the conditioner is outside the hot path
    betas = gd.linear_beta_schedule(K_STEP, max_beta=MAX_BETA)
    model = gd.GaussianDiffusion(None, 80, net, timesteps=K_STEP, K_step=K_STEP, loss_type="l1", betas=betas,
                                 spec_min=synth.SPEC_MIN, spec_max=synth.SPEC_MAX)
    return net, model


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    ns = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    gd = ns.gd
    net, model = build_reference_sampler(ns)
    out = {}
    with torch.no_grad():
        # schedule buffers (SURVEY.md §9.2 known-answer table comes from these)
        for k in ("betas", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped"):
            out["sched." + k] = getattr(model, k).numpy()
        # single DiffNet evaluations
        for i, c in enumerate(EPS_CASES):
            inp = synth.kernel_inputs(c["seed"], c["B"], c["T"], 1)
            eps = net(inp["start_noise"], torch.full((c["B"],), c["t"], dtype=torch.long), inp["cond"].transpose(1, 2))
            out[f"eps.{i}"] = eps.numpy()
        # full K-step sampler with injected noise
        for i, c in enumerate(DIFF_CASES):
            inp = synth.kernel_inputs(c["seed"], c["B"], c["T"], c["K"])
            noises = list(inp["step_noise"])
            gd.noise_like = lambda shape, device, repeat=False: noises.pop(0)
            xs = model.norm_spec(inp["fs2_mel"]).transpose(1, 2)[:, None]
            x = model.q_sample(xs, torch.tensor([K_STEP - 1]), noise=inp["start_noise"])
            for t in reversed(range(K_STEP)):
                x = model.p_sample(x, torch.full((c["B"],), t, dtype=torch.long), inp["cond"].transpose(1, 2))
            out[f"mel.{i}"] = model.denorm_spec(x[:, 0].transpose(1, 2)).numpy()
            out[f"x0.{i}"] = x.numpy()
        # vocoder
        import modules.parallel_wavegan.models.source as S  # type: ignore
        h = synth.HIFIGAN_CONFIG
        gen = ns.HifiGanGenerator(h)
        gen.remove_weight_norm()
        gen.load_state_dict(synth.hifigan_state(4321), strict=True)
        gen.eval()
        for i, c in enumerate(VOC_CASES):
            inp = synth.vocoder_inputs(c["seed"], c["B"], c["T"])

            class TorchProxy:   # replaces the name `torch` inside models/source.py for the RNG draws only
                def __getattr__(self, k):
                    return getattr(torch, k)

                def rand(self, *a, **k):
                    return inp["rand_ini"].clone()

                def randn_like(self, x):
                    return inp["src_noise"].clone() if x.shape[-1] == 9 else torch.zeros_like(x)

            S.torch = TorchProxy()
            f0_up = gen.f0_upsamp(inp["f0"][:, None]).transpose(1, 2)
            har, _, _ = gen.m_source(f0_up)
            out[f"har.{i}"] = har.transpose(1, 2).numpy()
            out[f"wav.{i}"] = gen(inp["mel"], inp["f0"]).numpy()
            out[f"wav_nof0.{i}"] = gen(inp["mel"]).numpy()
            S.torch = torch
    np.savez_compressed(os.path.join(OUT, "hotpath_golden.npz"), **out)
    sz = os.path.getsize(os.path.join(OUT, "hotpath_golden.npz"))
    print(f"wrote {len(out)} arrays, {sz / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
