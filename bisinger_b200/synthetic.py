"""Seeded synthetic weights and inputs for the hot path: what the tests, smoke() and bench.py feed the CUDA path and the oracle alike
(a data generator only -- no compute path of the package uses it; `oracle/synth.py` re-exports it for the oracle-side scripts).

No checkpoints or datasets are reachable, so parity tests and the bench use random-init
models of the named architecture (state-dict names/shapes per SURVEY.md §9.1) and synthetic
inputs.  Two reference defaults are deliberately NOT reproduced because they
make a parity test vacuous:
  * DiffNet.output_projection.weight is zero-initialised (usr/diff/net.py:105) -> eps would be
    the bias for any input; here it is kaiming-normal like the other convs (net.py:47-50);
  * HiFi-GAN res-block / upsample / post convs are N(0, 0.01) (modules/hifigan/hifigan.py:14-17)
    -> the MRF convolutions would contribute nothing; here every conv is N(0, 0.7*sqrt(2/fan_in)).
Everything is drawn from a CPU torch.Generator, so the same seed gives the same tensors in the
build container and on the GPU box (same image).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

Tensor = torch.Tensor

# usr/configs/lang-esm-style-ori-shift/base.yaml:76-77
SPEC_MAX = [-0.3894500136375427, -0.3796464204788208, -0.2914905250072479, -0.15550297498703003, -0.08502643555402756, 0.10698417574167252, -0.0739326998591423, -0.0541548952460289, 0.15501998364925385, 0.06483431905508041, 0.03054228238761425, -0.013737732544541359, -0.004876468330621719, 0.04368264228105545, 0.13329921662807465, 0.16471388936042786, 0.04605761915445328, -0.05680707097053528, 0.0542571023106575, -0.0076539707370102406, -0.00953489076346159, -0.04434828832745552, 0.001293870504014194, -0.12238839268684387, 0.06418416649103165, 0.02843189612030983, 0.08505241572856903, 0.07062800228595734, 0.00120724702719599, -0.07675088942050934, 0.03785804659128189, 0.04890783503651619, -0.06888376921415329, -0.0839693546295166, -0.17545585334300995, -0.2911079525947571, -0.4238220453262329, -0.262084037065506, -0.3002263605594635, -0.3845032751560211, -0.3906497061252594, -0.6550108790397644, -0.7810799479484558, -0.7503029704093933, -0.7995198965072632, -0.8092347383499146, -0.6196113228797913, -0.6684317588806152, -0.7735874056816101, -0.8324533104896545, -0.9601566791534424, -0.955253541469574, -0.748817503452301, -0.9106167554855347, -0.9707801342010498, -1.053107500076294, -1.0448424816131592, -1.1082794666290283, -1.1296544075012207, -1.071642279624939, -1.1003081798553467, -1.166810154914856, -1.1408926248550415, -1.1330615282058716, -1.1167492866516113, -1.0716774463653564, -1.035891056060791, -1.0092483758926392, -0.9675999879837036, -0.938962996006012, -1.0120564699172974, -0.9777995347976685, -1.029313564300537, -0.9459163546562195, -0.8519706130027771, -0.7751091122627258, -0.7933766841888428, -0.9019735455513, -0.9983296990394592, -1.505873441696167]
SPEC_MIN = [-6.0] * 80

# hop-128 HiFi-GAN/NSF layout (the config itself ships only with the
# external checkpoint, constraint from code: prod(upsample_rates) == hop_size == 128)
HIFIGAN_CONFIG = dict(
    resblock="1", upsample_rates=[8, 4, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4],
    upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    use_pitch_embed=True, audio_sample_rate=24000,
)

DIFFNET_CONFIG = dict(in_dims=80, hidden_size=256, residual_layers=20, residual_channels=256,
                      dilation_cycle_length=4)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _normal(g, shape, std):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def diffnet_state(seed: int = 1234, cfg: dict | None = None) -> Dict[str, Tensor]:
    """State dict with the parameter names of usr/diff/net.py:91-104."""
    c = dict(DIFFNET_CONFIG)
    c.update(cfg or {})
    M, H, C, L = c["in_dims"], c["hidden_size"], c["residual_channels"], c["residual_layers"]
    g = _gen(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, cout, cin, k):
        fan_in = cin * k
        sd[name + ".weight"] = _normal(g, (cout, cin, k), math.sqrt(2.0 / fan_in))  # kaiming_normal_
        sd[name + ".bias"] = _uniform(g, (cout,), 1.0 / math.sqrt(fan_in))

    def lin(name, cout, cin):
        b = 1.0 / math.sqrt(cin)
        sd[name + ".weight"] = _uniform(g, (cout, cin), b)
        sd[name + ".bias"] = _uniform(g, (cout,), b)

    conv("input_projection", C, M, 1)
    lin("mlp.0", 4 * C, C)
    lin("mlp.2", C, 4 * C)
    for i in range(L):
        p = f"residual_layers.{i}."
        conv(p + "dilated_conv", 2 * C, C, 3)
        lin(p + "diffusion_projection", C, C)
        conv(p + "conditioner_projection", 2 * C, H, 1)
        conv(p + "output_projection", 2 * C, C, 1)
    conv("skip_projection", C, C, 1)
    conv("output_projection", M, C, 1)
    return sd


def hifigan_state(seed: int = 4321, h: dict | None = None) -> Dict[str, Tensor]:
    """Weight-norm-FOLDED state dict (names as after remove_weight_norm(), hifigan.py:175-182)."""
    h = dict(HIFIGAN_CONFIG) if h is None else h
    g = _gen(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, cout, cin, k, fan_in=None):
        fan_in = fan_in or cin * k
        sd[name + ".weight"] = _normal(g, (cout, cin, k), 0.7 * math.sqrt(2.0 / fan_in))
        sd[name + ".bias"] = _uniform(g, (cout,), 1.0 / math.sqrt(fan_in))

    C0 = h["upsample_initial_channel"]
    sd["m_source.l_linear.weight"] = _uniform(g, (1, 9), 1.0 / 3.0)
    sd["m_source.l_linear.bias"] = _uniform(g, (1,), 1.0 / 3.0)
    conv("conv_pre", C0, 80, 7)
    rates, ks = h["upsample_rates"], h["upsample_kernel_sizes"]
    ch = C0
    for i, (u, k) in enumerate(zip(rates, ks)):
        cin, cout = C0 // (2 ** i), C0 // (2 ** (i + 1))
        # ConvTranspose1d weight is [in, out, k]; each output sample sees k/u taps of cin channels
        fan_in = cin * k // u
        sd[f"ups.{i}.weight"] = _normal(g, (cin, cout, k), 0.7 * math.sqrt(2.0 / fan_in))
        sd[f"ups.{i}.bias"] = _uniform(g, (cout,), 1.0 / math.sqrt(fan_in))
        if i + 1 < len(rates):
            s = int(np.prod(rates[i + 1:]))
            conv(f"noise_convs.{i}", cout, 1, 2 * s)
        else:
            conv(f"noise_convs.{i}", cout, 1, 1)
        ch = cout
    for i in range(len(rates)):
        c = C0 // (2 ** (i + 1))
        for j, k in enumerate(h["resblock_kernel_sizes"]):
            for m in range(len(h["resblock_dilation_sizes"][j])):
                conv(f"resblocks.{i * len(h['resblock_kernel_sizes']) + j}.convs1.{m}", c, c, k)
                conv(f"resblocks.{i * len(h['resblock_kernel_sizes']) + j}.convs2.{m}", c, c, k)
    conv("conv_post", 1, ch, 7)
    return sd


def pe_state(seed: int = 777, conv_layers: int = 2, n_mel: int = 80, C: int = 256) -> Dict[str, Tensor]:
    """State dict with the parameter/buffer names of PitchExtractor (modules/fastspeech/pe.py:120-136).  BatchNorm running
    statistics and every affine are non-trivial so that each term of the eval-mode formulas matters; the last Linear is
    scaled so that log2-f0 lands around 7..8.5 (130..360 Hz) with both signs of the uv logit."""
    g = _gen(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = _normal(g, (cout, cin, k), math.sqrt(2.0 / (cin * k)))
        sd[name + ".bias"] = _uniform(g, (cout,), 1.0 / math.sqrt(cin * k))

    def linear(name, cout, cin, gain=1.0):
        sd[name + ".weight"] = _uniform(g, (cout, cin), gain * math.sqrt(6.0 / (cin + cout)))
        sd[name + ".bias"] = _uniform(g, (cout,), 0.1)

    def affine(name, n):
        sd[name + ".weight"] = 1.0 + _uniform(g, (n,), 0.3)
        sd[name + ".bias"] = _uniform(g, (n,), 0.2)

    cin = n_mel
    for i in range(3):
        conv(f"mel_prenet.layers.{i}.0", C, cin, 5)
        affine(f"mel_prenet.layers.{i}.2", C)
        sd[f"mel_prenet.layers.{i}.2.running_mean"] = 0.3 + _uniform(g, (C,), 0.3)
        sd[f"mel_prenet.layers.{i}.2.running_var"] = 0.5 + torch.rand((C,), generator=g)
        sd[f"mel_prenet.layers.{i}.2.num_batches_tracked"] = torch.tensor(1000)
        cin = C
    linear("mel_prenet.out_proj", C, C)
    for j in range(conv_layers):
        conv(f"mel_encoder.conv.{j}.conv.conv", C, C, 5)
        affine(f"mel_encoder.conv.{j}.norm", C)
    if conv_layers > 0:
        linear("mel_encoder.in_proj", C, C)
        linear("mel_encoder.out_proj", C, C)
    sd["pitch_predictor.pos_embed_alpha"] = torch.tensor([0.8])
    for i in range(5):
        conv(f"pitch_predictor.conv.{i}.1", C, C, 5)
        affine(f"pitch_predictor.conv.{i}.3", C)
    sd["pitch_predictor.linear.weight"] = _uniform(g, (2, C), 0.04)
    sd["pitch_predictor.linear.bias"] = torch.tensor([7.8, 0.0])
    sd["pitch_predictor.embed_positions._float_tensor"] = torch.zeros(1)
    return sd


def fft_state(seed: int = 555, layers: int = 4, C: int = 256, k: int = 9, out_dims: int = 80) -> Dict[str, Tensor]:
    """State dict with the parameter / buffer names of FastspeechDecoder (modules/fastspeech/tts_modules.py:253-283,340-347;
    EncSALayer modules/commons/common_layers.py:680-701) plus ``mel_out.*`` (modules/fastspeech/fs2.py:60).  Non-trivial LayerNorm
    affines and a pos_embed_alpha != 1 so that every term matters."""
    g = _gen(seed)
    sd: Dict[str, Tensor] = {}
    sd["pos_embed_alpha"] = torch.tensor([0.9])
    sd["embed_positions._float_tensor"] = torch.zeros(1)
    xav = lambda out, inn: _uniform(g, (out, inn), math.sqrt(6.0 / (out + inn)))
    for i in range(layers):
        p = f"layers.{i}.op."
        for ln in ("layer_norm1", "layer_norm2"):
            sd[p + ln + ".weight"] = 1.0 + _uniform(g, (C,), 0.3)
            sd[p + ln + ".bias"] = _uniform(g, (C,), 0.2)
        sd[p + "self_attn.in_proj_weight"] = xav(3 * C, C) * 2.0        # sharper attention than xavier alone: the softmax must matter
        sd[p + "self_attn.out_proj.weight"] = xav(C, C)
        sd[p + "ffn.ffn_1.weight"] = _normal(g, (4 * C, C, k), math.sqrt(2.0 / (C * k)) * math.sqrt(k))
        sd[p + "ffn.ffn_1.bias"] = _uniform(g, (4 * C,), 0.1)
        sd[p + "ffn.ffn_2.weight"] = xav(C, 4 * C)
        sd[p + "ffn.ffn_2.bias"] = _uniform(g, (C,), 0.1)
    sd["layer_norm.weight"] = 1.0 + _uniform(g, (C,), 0.3)
    sd["layer_norm.bias"] = _uniform(g, (C,), 0.2)
    if out_dims:
        sd["mel_out.weight"] = xav(out_dims, C)
        sd["mel_out.bias"] = _uniform(g, (out_dims,), 0.5) - 3.0
    return sd


def fft_encoder_state(seed: int = 777, vocab: int = 62, layers: int = 4, C: int = 256, k: int = 9) -> Dict[str, Tensor]:
    """State dict with the names of FastspeechEncoder (modules/fastspeech/tts_modules.py:310-326: FFTBlocks(use_pos_embed=False) plus
    ``embed_tokens`` and its own ``embed_positions``): the decoder's layer names without ``pos_embed_alpha``, and
    ``embed_tokens.weight`` [vocab, C] ~ N(0, C^-0.5) with the padding row 0 zeroed (modules/commons/common_layers.py Embedding)."""
    sd = {k_: v for k_, v in fft_state(seed, layers, C, k, out_dims=0).items() if k_ != "pos_embed_alpha"}
    w = _normal(_gen(seed + 1), (vocab, C), C ** -0.5)
    w[0] = 0.0
    sd["embed_tokens.weight"] = w
    return sd


def fft_tokens(seed: int, B: int, T: int, vocab: int = 62, pad_tail: int = 0) -> Tensor:
    """txt_tokens [B,T] int64 in [1, vocab); the last ``pad_tail`` positions of every odd batch row are the padding index 0."""
    t = torch.randint(1, vocab, (B, T), generator=_gen(seed))
    if pad_tail > 0:
        t[1::2, T - pad_tail:] = 0
    return t


def fft_inputs(seed: int, B: int, T: int, C: int = 256, pad_tail: int = 0) -> Tensor:
    """decoder_inp [B,T,C] ~ N(0,1); the last ``pad_tail`` frames of every odd batch row are all-zero padding frames."""
    x = torch.randn((B, T, C), generator=_gen(seed))
    if pad_tail > 0:
        x[1::2, T - pad_tail:, :] = 0.0
    return x


def pe_inputs(seed: int, B: int, T: int, M: int = 80, pad_tail: int = 0) -> Tensor:
    """Log-mel input [B,T,80] as vocoder_inputs draws it; the last `pad_tail` frames of every odd batch row are all-zero
    padding frames (pe.py:30,145: padding = frames whose |mel| sums to 0)."""
    mel = vocoder_inputs(seed, B, T, M=M)["mel"].transpose(1, 2).contiguous()
    if pad_tail > 0:
        mel[1::2, T - pad_tail:, :] = 0.0
    return mel


def kernel_inputs(seed: int, B: int, T: int, K: int, M: int = 80, H: int = 256) -> Dict[str, Tensor]:
    """Kernel-level sampler inputs: cond ~ N(0,1) [B,T,H], fs2_mel in the log-mel
    range, q_sample noise and per-step noise z_k ~ N(0,1)."""
    g = _gen(seed)
    smin, smax = torch.tensor(SPEC_MIN[:M]), torch.tensor(SPEC_MAX[:M])
    return dict(
        cond=torch.randn((B, T, H), generator=g),
        fs2_mel=smin + torch.rand((B, T, M), generator=g) * (smax - smin),
        start_noise=torch.randn((B, 1, M, T), generator=g),
        step_noise=torch.randn((K, B, 1, M, T), generator=g),
    )


def vocoder_inputs(seed: int, B: int, T: int, hop: int = 128, M: int = 80) -> Dict[str, Tensor]:
    """Vocoder-only synthetic inputs: mel = spec_min + U(0,1)*(spec_max-spec_min)
    smoothed along t with a 9-tap box filter; f0 piece-wise constant per ~23-frame phoneme from
    MIDI 48..72 with every 7th phoneme unvoiced; NSF random phase and source noise."""
    g = _gen(seed)
    smin, smax = torch.tensor(SPEC_MIN[:M]), torch.tensor(SPEC_MAX[:M])
    u = torch.rand((B, T + 8, M), generator=g)
    u = torch.nn.functional.avg_pool1d(u.transpose(1, 2), 9, stride=1).transpose(1, 2)  # [B,T,M]
    mel = smin + u * (smax - smin)
    n_ph = max(1, (T * 8 * hop) // (24000 * 1) // 1)  # ~8 phonemes per second at 24 kHz
    n_ph = max(1, int(round(T * hop / 24000 * 8)))
    midi = torch.randint(48, 73, (B, n_ph), generator=g).float()
    hz = 440.0 * torch.pow(torch.tensor(2.0), (midi - 69.0) / 12.0)
    hz[:, 6::7] = 0.0
    base, extra = divmod(T, n_ph)
    reps = torch.tensor([base + (1 if i < extra else 0) for i in range(n_ph)])
    f0 = torch.repeat_interleave(hz, reps, dim=1)
    L = T * hop
    return dict(
        mel=mel.transpose(1, 2).contiguous(),       # [B,80,T] as run_vocoder hands it over
        f0=f0.contiguous(),                         # [B,T]
        rand_ini=torch.rand((B, 9), generator=g),
        src_noise=torch.randn((B, L, 9), generator=g),
    )
