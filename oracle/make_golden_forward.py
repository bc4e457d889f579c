"""Generate tests/golden/forward_golden.npz by executing the UNMODIFIED reference ``GaussianDiffusion.forward(infer=True)``
(TEST INFRASTRUCTURE ONLY; run in the build container:  python oracle/make_golden_forward.py).

The other sampler fixtures drive ``q_sample`` / ``p_sample`` directly; this one goes through the public entry point
(usr/diff/shallow_diffusion_tts.py:230-273) so that the handoff (``cond = decoder_inp.transpose``, ``norm_spec`` of the
FastSpeech2 mel, ``q_sample`` at t = K-1 or the Gaussian start, the loop selection by ``hparams['pndm_speedup']``, the
``mel2ph`` mask) is pinned too.  The FastSpeech2 conditioner is outside the hot path: ``fs2`` is a stub that returns the
synthetic ``decoder_inp`` / ``mel_out``.  Randomness: the reference draws from the global RNG, so every case runs under
``torch.manual_seed(rng_seed)`` and ``forward_noise()`` below re-draws the identical sequence for the oracle and the CUDA path
(CPU generator: same stream in the build container and on the GPU box, same image).

Cases:
  0  ancestral sampler, K = timesteps = 100, max_beta 0.06 (BASELINE's configuration), B = 2, T = 40, mel2ph with padding
  1  BiSinger's shipped configuration (usr/configs/lang-esm-style-ori-shift/diff.yaml:16-23): timesteps = K_step = 1000,
     max_beta 0.02, pndm_speedup 5, gaussian_start -> 200 PLMS iterations / 201 denoiser evaluations, B = 1 (the reference's
     p_sample_plms only runs for B = 1), T = 48
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
FWD_CASES = [
    dict(seed=71, rng_seed=171, B=2, T=40, timesteps=100, K_step=100, max_beta=0.06, pndm_speedup=None, gaussian_start=False,
         pad_tail=9),
    dict(seed=72, rng_seed=172, B=1, T=48, timesteps=1000, K_step=1000, max_beta=0.02, pndm_speedup=5, gaussian_start=True,
         pad_tail=0),
]


def forward_inputs(c):
    inp = synth.kernel_inputs(c["seed"], c["B"], c["T"], 1)
    mel2ph = torch.ones(c["B"], c["T"], dtype=torch.long)
    if c["pad_tail"]:
        mel2ph[-1, c["T"] - c["pad_tail"]:] = 0
    return inp["cond"], inp["fs2_mel"], mel2ph


def forward_noise(c):
    """The tensors ``forward(infer=True)`` draws under torch.manual_seed(rng_seed), in its order: randn_like of q_sample (:204,252),
    [randn of the Gaussian start (:256)], then one randn(x.shape) per ancestral step (:163; none with PLMS).
    Returns (start_noise [B,1,M,T], step_noise [K,B,1,M,T] or None)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(c["rng_seed"])
    shape = (c["B"], 1, 80, c["T"])
    # q_sample's randn_like(x_start): x_start is the permuted VIEW norm_spec(mel)[B,T,80].transpose(1,2)[:,None], and randn_like keeps
    # its strides -- the fill order follows that layout, so the draw is replayed on a tensor with the same strides
    start = torch.empty_like(torch.empty(c["B"], c["T"], 80).transpose(1, 2)[:, None]).normal_(generator=g).contiguous()
    if c["gaussian_start"]:
        start = torch.randn(shape, generator=g)
    if c["pndm_speedup"]:
        return start, None
    return start, torch.stack([torch.randn(shape, generator=g) for _ in range(c["K_step"])])


def main():
    import ref_shim
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    ns = ref_shim.load()
    gd = ns.gd
    gd.tqdm = lambda it, **k: it            # progress bar only
    out = {}
    sd = synth.diffnet_state(1234)
    for i, c in enumerate(FWD_CASES):
        ns.hparams.update(pndm_speedup=c["pndm_speedup"], gaussian_start=c["gaussian_start"], timesteps=c["timesteps"],
                          K_step=c["K_step"], max_beta=c["max_beta"])
        cond, fs2_mel, mel2ph = forward_inputs(c)

        class StubFs2(torch.nn.Module):     # the conditioner is outside the hot path
            def forward(self, txt_tokens, mel2ph_, spk_embed, ref_mels, f0, uv, energy, skip_decoder=False, infer=True, **kw):
                return {"decoder_inp": cond, "mel_out": fs2_mel}

        gd.FastSpeech2 = lambda *a, **k: StubFs2()
        net = ns.DiffNet(80)
        net.load_state_dict(sd, strict=True)
        net.eval()
        model = gd.GaussianDiffusion(None, 80, net, timesteps=c["timesteps"], K_step=c["K_step"], loss_type="l1",
                                     betas=gd.linear_beta_schedule(c["timesteps"], max_beta=c["max_beta"]),
                                     spec_min=synth.SPEC_MIN, spec_max=synth.SPEC_MAX)
        torch.manual_seed(c["rng_seed"])
        with torch.no_grad():
            ret = model(torch.zeros(c["B"], 8, dtype=torch.long), mel2ph=mel2ph, infer=True)
        out[f"mel.{i}"] = ret["mel_out"].numpy()
        assert torch.equal(ret["fs2_mel"], fs2_mel)
        print(f"case {i}: mel_out {tuple(ret['mel_out'].shape)} range [{float(ret['mel_out'].min()):.3f}, {float(ret['mel_out'].max()):.3f}]")
    ns.hparams.update(ref_shim.HOTPATH_HPARAMS)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "forward_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
