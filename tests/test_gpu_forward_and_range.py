"""GPU tests (-m gpu): (1) the drop-in's public entry point ``B200GaussianDiffusion.forward(infer=True)`` against fixtures written by
the executed reference's ``GaussianDiffusion.forward(infer=True)`` (oracle/make_golden_forward.py) -- the K = 100 ancestral sampler and
BiSinger's shipped PLMS configuration (timesteps = K_step = 1000, max_beta 0.02, pndm_speedup 5, gaussian_start); (2) the PLMS loop
under a CUDA graph; (3) range robustness of the fp16x2 contraction (large conditioner inputs, outlier weight rows, mels at the
spec_min / spec_max edges) and the automatic, tested fall-back to bf16x3 when the weights' worst-case activation bound leaves the
fp16 range.  Tolerance: BASELINE.json north_star (mel max-abs <= 1e-2), scaled to the reference's output range where a random-init
denoiser drives the un-clamped PLMS trajectory out of the mel range (see _fwd_tol)."""
import os

import numpy as np
import pytest
import torch

import svs_oracle as O
import synth
from make_golden import K_STEP, MAX_BETA
from make_golden_forward import FWD_CASES, forward_inputs, forward_noise

pytestmark = pytest.mark.gpu
MEL_TOL = 1e-2
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "forward_golden.npz")


def _fwd_tol(ref):
    return MEL_TOL * max(1.0, float(np.abs(np.asarray(ref)).max()) / 6.2)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def _model(dev, c, cond, fs2_mel, sd=None, precision="fp16x2"):
    from bisinger_b200 import B200DiffNet, B200GaussianDiffusion
    net = B200DiffNet(80)
    net.load_state_dict(sd or synth.diffnet_state(1234), strict=True)

    class StubFs2(torch.nn.Module):       # the FastSpeech2 conditioner is outside the hot path
        def forward(self, txt_tokens, mel2ph, spk_embed, ref_mels, f0, uv, energy, skip_decoder=False, infer=True, **kw):
            return {"decoder_inp": cond.to(dev), "mel_out": fs2_mel.to(dev)}

    hp = dict(hidden_size=256, residual_layers=20, residual_channels=256, dilation_cycle_length=4, keep_bins=80,
              pndm_speedup=c["pndm_speedup"], gaussian_start=c["gaussian_start"])
    return B200GaussianDiffusion(None, 80, net, timesteps=c["timesteps"], K_step=c["K_step"],
                                 betas=O.linear_beta_schedule(c["timesteps"], c["max_beta"]), spec_min=synth.SPEC_MIN,
                                 spec_max=synth.SPEC_MAX, fs2=StubFs2(), hparams=hp, precision=precision).to(dev)


@pytest.mark.parametrize("i", range(len(FWD_CASES)))
def test_forward_infer_vs_reference_forward_golden(dev, i):
    c = FWD_CASES[i]
    g = np.load(GOLDEN)
    cond, fs2_mel, mel2ph = forward_inputs(c)
    start, steps = forward_noise(c)
    model = _model(dev, c, cond, fs2_mel)
    ret = model(torch.zeros(c["B"], 8, dtype=torch.long, device=dev), mel2ph=mel2ph.to(dev), infer=True, start_noise=start, step_noise=steps)
    assert {"mel_out", "fs2_mel", "decoder_inp"} <= set(ret)
    ref = g[f"mel.{i}"]
    err = float(np.abs(ret["mel_out"].cpu().numpy() - ref).max())
    assert err <= _fwd_tol(ref), f"case {i}: {err} > {_fwd_tol(ref)} (reference range {np.abs(ref).max():.1f})"
    if c["pad_tail"]:
        assert float(ret["mel_out"][-1, c["T"] - c["pad_tail"]:].abs().max()) == 0.0


def test_plms_shipped_config_batch_vs_oracle_and_graph_replay(dev, monkeypatch):
    """The shipped configuration for B > 1 (the reference's own loop only runs for B = 1) against the oracle, twice through the
    captured graph (bit-identical replays), and against plain launches (BSG_DIFF_GRAPH=0: bit-identical)."""
    c = dict(FWD_CASES[1], B=3, T=70, seed=73, rng_seed=173)
    cond, fs2_mel, mel2ph = forward_inputs(c)
    mel2ph[1, 50:] = 0
    start, _ = forward_noise(c)
    sd = synth.diffnet_state(1234)
    sched = O.schedule_buffers(O.linear_beta_schedule(c["timesteps"], c["max_beta"]))
    with torch.no_grad():
        ref = O.diffusion_infer_plms(sd, sched, torch.tensor(synth.SPEC_MIN), torch.tensor(synth.SPEC_MAX), cond, c["K_step"],
                                     c["pndm_speedup"], fs2_mel, start, mel2ph=mel2ph, gaussian_start=True)
    model = _model(dev, c, cond, fs2_mel)
    tok = torch.zeros(c["B"], 8, dtype=torch.long, device=dev)
    a = model(tok, mel2ph=mel2ph.to(dev), infer=True, start_noise=start)["mel_out"]
    b = model(tok, mel2ph=mel2ph.to(dev), infer=True, start_noise=start)["mel_out"]
    assert torch.equal(a, b)
    err = float((a.cpu() - ref).abs().max())
    assert err <= _fwd_tol(ref.numpy()), f"{err} > {_fwd_tol(ref.numpy())}"
    assert float(a[1, 50:].abs().max()) == 0.0
    monkeypatch.setenv("BSG_DIFF_GRAPH", "0")
    plain = _model(dev, c, cond, fs2_mel)
    p = plain(tok, mel2ph=mel2ph.to(dev), infer=True, start_noise=start)["mel_out"]
    assert torch.equal(a, p)


def _run(dev, sd, inp, precision="fp16x2"):
    from bisinger_b200 import B200DiffNet, DiffusionPlan
    net = B200DiffNet(80)
    net.load_state_dict(sd, strict=True)
    sched = O.schedule_buffers(O.linear_beta_schedule(K_STEP, MAX_BETA))
    plan = DiffusionPlan(net, sched, K_STEP, K_STEP, synth.SPEC_MIN, synth.SPEC_MAX, precision=precision, device=dev)
    mel = plan.sample(inp["cond"].to(dev), inp["fs2_mel"].to(dev), inp["start_noise"].to(dev), inp["step_noise"].to(dev)).cpu()
    with torch.no_grad():
        ref = O.diffusion_infer(sd, sched, torch.tensor(synth.SPEC_MIN), torch.tensor(synth.SPEC_MAX), inp["cond"], K_STEP,
                                inp["step_noise"], inp["fs2_mel"], inp["start_noise"])
    return plan, mel, ref


def test_range_large_conditioner_and_edge_mels(dev):
    """cond x 30 (gate pre-activations deep in saturation, conditioner projection ~ +-60) and a FastSpeech2 mel sitting on
    spec_min / spec_max (x_start = +-1 exactly, the clamp edge)."""
    B, T = 2, 300
    inp = synth.kernel_inputs(501, B, T, K_STEP)
    inp["cond"] = inp["cond"] * 30.0
    smin, smax = torch.tensor(synth.SPEC_MIN), torch.tensor(synth.SPEC_MAX)
    inp["fs2_mel"][0, ::2] = smin
    inp["fs2_mel"][0, 1::2] = smax
    plan, mel, ref = _run(dev, synth.diffnet_state(1234), inp)
    assert plan.precision == "fp16x2"
    assert float((mel - ref).abs().max()) <= MEL_TOL


def _with_outliers(scale, names=("dilated_conv", "conditioner_projection", "output_projection")):
    sd = synth.diffnet_state(1234)
    g = torch.Generator().manual_seed(9)
    for l in (0, 7, 19):
        for name in names:
            rows = torch.randint(0, 512, (3,), generator=g)
            if name == "output_projection":
                rows = rows % 256 + 256          # skip half (the residual half is bounded separately by fp16_activation_bound)
            sd[f"residual_layers.{l}.{name}.weight"][rows] *= scale
    return sd


def test_range_outlier_weight_rows(dev):
    """A few outlier rows in the dilated conv / conditioner / output projections of several layers.  The fp16 weight packing is scaled
    by the largest weight of each matrix (the ordinary weights lose head-room) and the skip path multiplies the 11-bit rounding of the
    gated activations by its gain.  10x outliers (and 100x outliers off the skip path) keep fp16x2 and must meet the tolerance; 100x
    outliers in the skip halves push the skip-path gain past FP16_SKIP_GAIN_MAX (measured: 8.5e-3 .. 2e-2 in fp16x2,
    tests/tools/exp_outlier.py), so the plan falls back to bf16x3 -- and meets the tolerance there."""
    from bisinger_b200.diffusion import FP16_SAFE_BOUND, FP16_SKIP_GAIN_MAX, fp16_activation_bound, fp16_skip_gain, B200DiffNet
    inp = synth.kernel_inputs(502, 2, 200, K_STEP)
    for sd, want in ((_with_outliers(10.0), "fp16x2"), (_with_outliers(100.0, ("dilated_conv", "conditioner_projection")), "fp16x2"),
                     (_with_outliers(100.0), "bf16x3")):
        net = B200DiffNet(80)
        net.load_state_dict(sd, strict=True)
        assert fp16_activation_bound(net) < FP16_SAFE_BOUND
        assert (fp16_skip_gain(net) <= FP16_SKIP_GAIN_MAX) == (want == "fp16x2")
        if want == "bf16x3":
            with pytest.warns(UserWarning, match="skip-path gain"):
                plan, mel, ref = _run(dev, sd, inp)
        else:
            plan, mel, ref = _run(dev, sd, inp)
        assert plan.precision == want
        assert float((mel - ref).abs().max()) <= MEL_TOL, want


def test_range_guard_falls_back_to_bf16x3(dev):
    """Residual-path weights whose worst-case bound leaves the fp16 range (x1000: |x| may reach 5e4 > 65504 / 2): the plan is built in
    bf16x3 (fp32 exponent range) with a warning.  With weights like these the network is chaotic (gates flip between saturation
    levels on 1-ulp differences; even fp32 carries only ~1e-2 absolute precision at |x| ~ 1e4..1e5), so no implementation can be held
    to a tolerance against the oracle here: the test checks the DECISION and that the output is finite and inside the mel range (the
    x0 clamp).  The tolerance-level fall-back test is test_range_outlier_weight_rows.  Forcing fp16x2 stays finite, too (saturating
    conversions, never inf)."""
    from bisinger_b200 import B200DiffNet
    from bisinger_b200.diffusion import FP16_SAFE_BOUND, fp16_activation_bound
    sd = synth.diffnet_state(1234)
    for l in range(20):
        sd[f"residual_layers.{l}.output_projection.weight"][:256] *= 1000.0
    net = B200DiffNet(80)
    net.load_state_dict(sd, strict=True)
    assert fp16_activation_bound(net) >= FP16_SAFE_BOUND
    inp = synth.kernel_inputs(503, 2, 130, K_STEP)
    with pytest.warns(UserWarning, match="bf16x3"):
        plan, mel, ref = _run(dev, sd, inp)
    assert plan.precision == "bf16x3" and "activation bound" in plan.precision_note
    smin, smax = torch.tensor(synth.SPEC_MIN), torch.tensor(synth.SPEC_MAX)
    assert bool(torch.isfinite(mel).all()) and bool((mel >= smin - 1e-4).all()) and bool((mel <= smax + 1e-4).all())
    forced, mel16, _ = _run(dev, sd, inp, precision="fp16x2!")
    assert forced.precision == "fp16x2" and bool(torch.isfinite(mel16).all())
