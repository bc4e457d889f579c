"""GPU bring-up probe: runs each stage in its own subprocess (a kernel trap poisons the CUDA context) with a
timeout, prints diagnostics instead of asserting.  Usage: python tests/tools/gpu_probe.py [stage ...]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def conv_ref(a, w, bias, shifts):
    import torch
    B, L, Cin = a.shape
    N = w.shape[0]
    out = bias.double()[None, None, :].repeat(B, L, 1)
    for tp, s in enumerate(shifts):
        sh = torch.zeros_like(a, dtype=torch.float64)
        lo, hi = max(0, -s), min(L, L - s)
        if hi > lo:
            sh[:, lo:hi] = a[:, lo + s:hi + s].double()
        out += sh @ w[:, tp, :].double().t()
    return out


def stage_selftest():
    import ctypes as C, torch
    from bisinger_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(0)
    cases = [
        # B, L, Cin, N, shifts, n_tile, precision
        (1, 128, 64, 128, [0], 128, 0),
        (1, 128, 64, 256, [0], 256, 0),
        (1, 128, 256, 256, [0], 256, 0),
        (2, 300, 256, 512, [-2, 0, 2], 256, 0),
        (2, 300, 256, 512, [-8, 0, 8], 256, 1),
        (1, 128, 64, 128, [0], 128, 1),
        (3, 77, 80, 256, [0], 256, 1),
        (2, 200, 64, 64, [-1, 0, 1], 64, 0),
        (2, 200, 64, 32, [-3, 0, 3], 32, 0),
        (1, 1000, 128, 128, [-5, -4, -3, -2, -1, 0, 1, 2, 3, 4, 5], 128, 0),
    ]
    for (B, Lr, Cin, N, shifts, n_tile, prec) in cases:
        a = torch.randn(B, Lr, Cin, device="cuda")
        w = torch.randn(N, len(shifts), Cin) / (Cin * len(shifts)) ** 0.5
        bias = torch.randn(N)
        out = torch.full((B, Lr, N), float("nan"), device="cuda")
        sh = (C.c_int * len(shifts))(*shifts)
        st = L.bsg_selftest_conv(_lib.dev_ptr(a), _lib.fptr(w.contiguous()), _lib.fptr(bias), B, Lr, Cin, N, len(shifts), sh, n_tile, prec,
                                 _lib.dev_ptr(out), None)
        if st != 0:
            print("CASE", (B, Lr, Cin, N, shifts, n_tile, prec), "ERROR", L.bsg_last_error().decode()); continue
        torch.cuda.synchronize()
        if prec == 0:
            ar, wr = a.bfloat16().float(), w.bfloat16().float()
        else:
            ar, wr = a, w
        ref = conv_ref(ar, wr.cuda(), bias.cuda(), shifts)
        err = (out.double() - ref).abs()
        print("CASE", (B, Lr, Cin, N, shifts, n_tile, prec), "max_err %.3e" % err.max().item(), "ref_rms %.3f" % ref.pow(2).mean().sqrt().item(),
              "nan", int(torch.isnan(out).sum().item()))
        if not (err.max().item() < 1e-2):
            e = err[0]
            rows = e.amax(dim=1)
            cols = e.amax(dim=0)
            print("   bad rows (first 40 of b=0):", [i for i in range(min(Lr, 400)) if rows[i] > 1e-2][:40])
            print("   bad cols (first 40):", [i for i in range(N) if cols[i] > 1e-2][:40])
            print("   out[0,0,:8]", out[0, 0, :8].tolist(), "ref", ref[0, 0, :8].tolist())


def _diff_setup(B, T, K, prec):
    import torch
    import synth, svs_oracle as O
    from bisinger_b200 import B200DiffNet, DiffusionPlan
    sd = synth.diffnet_state(1234)
    net = B200DiffNet(80)
    net.load_state_dict(sd, strict=True)
    sched = O.schedule_buffers(O.linear_beta_schedule(K, 0.06))
    plan = DiffusionPlan(net, sched, K, K, synth.SPEC_MIN, synth.SPEC_MAX, precision=prec, device=torch.device("cuda", 0))
    inp = synth.kernel_inputs(7, B, T, K)
    return sd, sched, plan, inp, O, synth


def stage_denoise():
    import torch
    for prec in ("bf16x3", "bf16"):
        B, T, K = 2, 200, 100
        sd, sched, plan, inp, O, synth = _diff_setup(B, T, K, prec)
        x = inp["start_noise"]
        for t in (99, 50, 0):
            ref = O.diffnet_forward(sd, x, torch.full((B,), t), inp["cond"].transpose(1, 2))
            emu = O.diffnet_forward(sd, x, torch.full((B,), t), inp["cond"].transpose(1, 2), operand="bf16")
            out = plan.denoise(x.cuda(), t, inp["cond"].cuda()).cpu()
            print("DENOISE", prec, "t", t, "max|out-ref| %.3e" % (out - ref).abs().max().item(), "max|emu_bf16-ref| %.3e" % (emu - ref).abs().max().item(),
                  "ref rms %.3f" % ref.pow(2).mean().sqrt().item(), "nan", int(torch.isnan(out).sum()))


def stage_sample():
    import torch
    cases = (("bf16x3", (2, 100)), ("bf16x3", (1, 333)), ("bf16", (2, 100)))
    if os.environ.get("BSG_PREC"):
        cases = tuple((os.environ["BSG_PREC"], bt) for bt in ((2, 100), (1, 333), (3, 700)))
    for prec, (B, T) in cases:
        K = 100
        sd, sched, plan, inp, O, synth = _diff_setup(B, T, K, prec)
        smin, smax = torch.tensor(synth.SPEC_MIN), torch.tensor(synth.SPEC_MAX)
        t0 = time.time()
        ref = O.diffusion_infer(sd, sched, smin, smax, inp["cond"], K, inp["step_noise"], inp["fs2_mel"], inp["start_noise"])
        t1 = time.time()
        mel = plan.sample(inp["cond"].cuda(), inp["fs2_mel"].cuda(), inp["start_noise"].cuda(), inp["step_noise"].cuda()).cpu()
        torch.cuda.synchronize()
        print("SAMPLE", prec, (B, T), "max|mel-ref| %.3e" % (mel - ref).abs().max().item(), "mean %.3e" % (mel - ref).abs().mean().item(),
              "oracle_s %.2f" % (t1 - t0), "nan", int(torch.isnan(mel).sum()))
        # graph path (device RNG): just check it runs, is finite and deterministic for a seed
        m1 = plan.sample(inp["cond"].cuda(), inp["fs2_mel"].cuda(), None, None, seed=5)
        m2 = plan.sample(inp["cond"].cuda(), inp["fs2_mel"].cuda(), None, None, seed=5)
        m3 = plan.sample(inp["cond"].cuda(), inp["fs2_mel"].cuda(), None, None, seed=6)
        torch.cuda.synchronize()
        print("   graph path: finite", bool(torch.isfinite(m1).all()), "same-seed equal", bool(torch.equal(m1, m2)), "diff-seed differs",
              bool(not torch.equal(m1, m3)), "mel range", m1.min().item(), m1.max().item())


def stage_bench():
    import torch
    from bisinger_b200 import _lib
    for prec in ("bf16x3", "bf16"):
        for (B, T) in ((1, 938), (32, 1875)):
            K = 100
            sd, sched, plan, inp, O, synth = _diff_setup(1, 8, K, prec)
            g = torch.Generator().manual_seed(1)
            cond = torch.randn(B, T, 256, generator=g).cuda()
            fs2 = (torch.rand(B, T, 80, generator=g) * 5 - 6).cuda()
            for i in range(2):
                plan.sample(cond, fs2, None, None, seed=i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = _lib.launch_count()
            e0.record()
            for i in range(3):
                plan.sample(cond, fs2, None, None, seed=10 + i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            flops = 26427392.0 * B * T * K
            print("BENCH", prec, (B, T), "ms/sample %.2f" % ms, "audio_s/s %.1f" % (B * T / 187.5 / (ms / 1e3)),
                  "algorithmic TFLOP/s %.1f" % (flops / ms / 1e9), "launches/sample", (_lib.launch_count() - n0) // 3)


def _voc_setup():
    import torch
    import synth, svs_oracle as O
    from bisinger_b200.vocoder import B200HifiGanGenerator
    h = synth.HIFIGAN_CONFIG
    sd = synth.hifigan_state(4321)
    gen = B200HifiGanGenerator(h)
    gen.load_folded_state_dict(sd, strict=True)
    gen.build_plan(torch.device("cuda", 0))
    return h, sd, gen, O, synth


def stage_vocoder():
    import torch
    h, sd, gen, O, synth = _voc_setup()
    for (B, T) in ((2, 64), (1, 300), (3, 131)):
        inp = synth.vocoder_inputs(11, B, T)
        t0 = time.time()
        ref, har_ref = O.hifigan_forward(sd, h, inp["mel"], inp["f0"], inp["rand_ini"], inp["src_noise"], return_source=True)
        emu = O.hifigan_forward(sd, h, inp["mel"], inp["f0"], inp["rand_ini"], inp["src_noise"], operand="bf16")
        t1 = time.time()
        har = gen.plan.source(inp["f0"], inp["rand_ini"], inp["src_noise"]).cpu()
        print("SOURCE", (B, T), "max|har-ref| %.3e" % (har - har_ref[:, 0]).abs().max().item(), "har rms %.3f" % har_ref.pow(2).mean().sqrt().item())
        out = gen(inp["mel"].cuda(), inp["f0"].cuda(), inp["rand_ini"].cuda(), inp["src_noise"].cuda()).cpu()
        print("VOCODER", (B, T), "SNR dB %.2f" % O.snr_db(ref, out), "emulated-bf16 SNR %.2f" % O.snr_db(ref, emu), "SNR vs emu %.2f" % O.snr_db(emu, out),
              "max abs %.3e" % (out - ref).abs().max().item(), "oracle_s %.2f" % (t1 - t0), "nan", int(torch.isnan(out).sum()))
        out2 = gen(inp["mel"].cuda(), None).cpu()
        ref2 = O.hifigan_forward(sd, h, inp["mel"], None)
        print("   no-f0 path SNR dB %.2f" % O.snr_db(ref2, out2))
        out3 = gen(inp["mel"].cuda(), inp["f0"].cuda(), seed=3)
        print("   device-RNG path finite", bool(torch.isfinite(out3).all()), "rms %.3f" % out3.pow(2).mean().sqrt().item())


def stage_vbench():
    import torch
    from bisinger_b200 import _lib
    h, sd, gen, O, synth = _voc_setup()
    for (B, T) in ((1, 938), (32, 1875), (8, 11250)):
        g = torch.Generator().manual_seed(1)
        mel = (torch.rand(B, 80, T, generator=g) * 5 - 6).cuda()
        f0 = (torch.rand(B, T, generator=g) * 300 + 100).cuda()
        for i in range(2):
            gen(mel, f0, seed=i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for i in range(3):
            gen(mel, f0, seed=10 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        flops = 375734272.0 * B * T
        print("VBENCH", (B, T), "ms %.2f" % ms, "audio_s/s %.1f" % (B * T / 187.5 / (ms / 1e3)), "algorithmic TFLOP/s %.1f" % (flops / ms / 1e9),
              "launches", (_lib.launch_count() - n0) // 3, "mem GB %.1f" % (torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9))


def stage_proftarget():
    """Target for ncu: a few launches of one DiffNet layer kernel (BSG_WHICH = 0 gate / 1 residual-skip) at the cfg3 shape."""
    import torch
    sd, sched, plan, inp, O, synth = _diff_setup(1, 8, 100, os.environ.get("BSG_PREC", "fp16x2"))
    which = int(os.environ.get("BSG_WHICH", "1"))
    print("kernel", which, "ms", plan.time_kernel(which, 32, 1875, int(os.environ.get("BSG_REPS", "4"))))


def stage_mctest():
    """two CTA pairs per cluster with multicast weights (precision | 0x200)"""
    stage_pairtest(0x200)


def stage_pairtest(flag=0x100):
    """2-CTA (cta_group::2) variant of the conv kernel through the self-test entry (precision | 0x100)."""
    import ctypes as C, torch
    from bisinger_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(0)
    cases = [
        (1, 256, 64, 256, [0], 256, 0), (1, 256, 256, 256, [0], 256, 0), (2, 300, 256, 512, [-2, 0, 2], 256, 0),
        (2, 300, 256, 512, [-8, 0, 8], 256, 1), (1, 1875, 256, 256, [0], 128, 1), (3, 77, 256, 256, [-1, 0, 1], 256, 1),
        (4, 1000, 512, 256, [0], 128, 1), (5, 600, 256, 512, [-4, 0, 4], 256, 2), (32, 1875, 256, 512, [-8, 0, 8], 256, 2),
    ]
    if flag == 0x200:
        cases = [c for c in cases if c[5] == 256]
    for (B, Lr, Cin, N, shifts, n_tile, prec) in cases:
        a = torch.randn(B, Lr, Cin, device="cuda")
        w = torch.randn(N, len(shifts), Cin) / (Cin * len(shifts)) ** 0.5
        bias = torch.randn(N)
        out = torch.full((B, Lr, N), float("nan"), device="cuda")
        sh = (C.c_int * len(shifts))(*shifts)
        st = L.bsg_selftest_conv(_lib.dev_ptr(a), _lib.fptr(w.contiguous()), _lib.fptr(bias), B, Lr, Cin, N, len(shifts), sh, n_tile, prec | flag,
                                 _lib.dev_ptr(out), None)
        if st != 0:
            print("PAIR CASE", (B, Lr, Cin, N, shifts, n_tile, prec), "ERROR", L.bsg_last_error().decode()); continue
        torch.cuda.synchronize()
        ar, wr = (a.bfloat16().float(), w.bfloat16().float()) if prec == 0 else (a, w)
        ref = conv_ref(ar, wr.cuda(), bias.cuda(), shifts)
        err = (out.double() - ref).abs()
        print("PAIR CASE", (B, Lr, Cin, N, shifts, n_tile, prec), "max_err %.3e" % err.max().item(), "nan", int(torch.isnan(out).sum().item()))
        if not (err.max().item() < 1e-2):
            e = err[0]
            rows = e.amax(dim=1); cols = e.amax(dim=0)
            print("   bad rows (b=0):", [i for i in range(min(Lr, 600)) if rows[i] > 1e-2][:24], "... count", int((rows > 1e-2).sum()))
            print("   bad cols:", [i for i in range(N) if cols[i] > 1e-2][:24], "... count", int((cols > 1e-2).sum()))


def stage_vproftarget():
    """Target for ncu: two vocoder forwards at the cfg3 shape."""
    import torch
    h, sd, gen, O, synth = _voc_setup()
    B, T = 32, 1875
    g = torch.Generator().manual_seed(1)
    mel = (torch.rand(B, 80, T, generator=g) * 5 - 6).cuda()
    f0 = (torch.rand(B, T, generator=g) * 300 + 100).cuda()
    for i in range(2):
        gen(mel, f0, seed=i)
    torch.cuda.synchronize()


STAGES = {"selftest": stage_selftest, "denoise": stage_denoise, "sample": stage_sample, "bench": stage_bench, "vocoder": stage_vocoder, "vbench": stage_vbench, "proftarget": stage_proftarget, "pairtest": stage_pairtest, "mctest": stage_mctest, "vproftarget": stage_vproftarget}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    stages = sys.argv[1:] or list(STAGES)
    for s in stages:
        print(f"===== stage {s} =====", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", s], timeout=420, capture_output=True, text=True)
            print(r.stdout[-8000:])
            if r.returncode != 0:
                print("RC", r.returncode, "STDERR tail:\n", r.stderr[-3000:])
        except subprocess.TimeoutExpired as e:
            print("TIMEOUT", s, (e.stdout or b"")[-3000:] if e.stdout else "")
        print(f"----- {s} took {time.time() - t0:.1f}s", flush=True)
