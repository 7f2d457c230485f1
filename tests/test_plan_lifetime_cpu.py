"""Buffer lifetimes of the recorded launch plans (tools/plan_check.py), on the CPU.

Pointers are baked into a plan when it is recorded and activations come out of an exact-size free-list arena
(unet_engine.Arena), so a tensor released one operator too early corrupts results only at the batch sizes where a later
allocation happens to have the same rounded size -- the GPU parity tests run a handful of shapes.  Recording needs no
device, so every plan variant is recorded here against a recorder in place of the library and checked operator by
operator: operands live when recorded, every arena buffer read (``const`` in include/sonic.h) written since it was
handed out (the cached DeepCache plan reads only what the full plan left resident), no double release, and the bytes
each operator touches (rows x pitch of every GEMM / attention / norm operand, the partial-statistics tables) inside the
arena buffer they start in.
"""
import ctypes as C
import os
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)

import plan_check  # noqa: E402


def test_checker_detects_injected_faults():
    from sonicdiffusionbayeslab_b200 import kernels as K
    from sonicdiffusionbayeslab_b200 import unet_engine as UE

    with plan_check.recording() as tr:
        lib = UE.lib()
        arena, plan = UE.Arena(torch.device("cpu")), UE._Plan()
        g = torch.ones(64)

        def layernorm(x, y):
            assert lib.sonic_plan_add_layernorm(plan.h, K.ptr(x), K.ptr(y), 64, 64, C.c_float(1e-5), K.ptr(g), K.ptr(g)) == 0

        a, b, c = arena.alloc((64, 64)), arena.alloc((64, 64)), arena.alloc((64, 64))
        x = torch.zeros(64, 64, dtype=torch.bfloat16)                 # not an arena buffer: an engine input
        layernorm(x, a)
        layernorm(a, b)
        assert tr.problems == [] and tr.arena_reads == 1 and tr.arena_writes == 2
        layernorm(c, b)                                               # c was never written
        assert len(tr.problems) == 1 and "nobody has written" in tr.problems[-1]
        arena.release(a)
        layernorm(a, b)                                               # use after release
        assert len(tr.problems) == 2 and "RELEASED" in tr.problems[-1]
        d = arena.alloc((64, 64))                                     # recycles a's buffer ...
        assert d.data_ptr() == a.data_ptr()
        layernorm(a, b)                                               # ... so the stale pointer reads an unwritten buffer
        assert len(tr.problems) == 3 and "nobody has written" in tr.problems[-1]
        arena.release(b)
        arena.release(b)
        assert len(tr.problems) == 4 and "released twice" in tr.problems[-1]
        e, f = arena.alloc((64, 64)), arena.alloc((128, 64))
        layernorm(x, e)
        n = len(tr.problems)
        assert lib.sonic_plan_add_layernorm(plan.h, K.ptr(e), K.ptr(f), 128, 64, C.c_float(1e-5), K.ptr(g), K.ptr(g)) == 0
        assert len(tr.problems) == n + 1 and "its arena buffer ends" in tr.problems[-1]   # 128 rows read from a 64-row buffer
        # a GEMM writing over its own input, and an operand off the 16-byte grid
        from sonicdiffusionbayeslab_b200._lib import GemmArgs

        w = torch.zeros(64, 64, dtype=torch.bfloat16)

        def gemm(a0, out, ld_out=64):
            ga = GemmArgs()
            ga.a0, ga.c0, ga.ld0 = a0.data_ptr(), 64, 64
            ga.n_img, ga.H, ga.W, ga.w, ga.N, ga.taps = 1, 1, 64, w.data_ptr(), 64, 1
            ga.out, ga.ld_out = out.data_ptr(), ld_out
            assert lib.sonic_plan_add_conv_gemm(plan.h, C.byref(ga)) == 0

        n = len(tr.problems)
        gemm(e, f)
        assert len(tr.problems) == n
        gemm(e, e)
        assert len(tr.problems) == n + 1 and "overlaps input" in tr.problems[-1]
        gemm(e, f.view(-1)[4:].view(-1), ld_out=64)                   # starts 8 bytes into the buffer
        assert len(tr.problems) == n + 2 and "TMA-addressable" in tr.problems[-1]
        gemm(e, f, ld_out=68)                                         # 136-byte pitch
        assert "TMA-addressable" in tr.problems[-1] and len(tr.problems) >= n + 3
    assert UE.lib is not None and UE.Arena.alloc.__name__ == "alloc" and not torch.zeros(1).is_cuda   # hooks restored


@pytest.fixture(scope="module")
def packed():
    from sonicdiffusionbayeslab_b200.unet_engine import PackedWeights
    from sonicdiffusionbayeslab_b200.unet_spec import random_unet_state_dict

    return PackedWeights(random_unet_state_dict(29), "cpu")


@pytest.mark.parametrize("n_latents,cfg", [(1, True), (16, True), (16, False), (32, True), (64, False)])
def test_unet_plans_have_no_lifetime_hazards(packed, n_latents, cfg):
    """ctx + full + cached plan of every DeepCache branch 0-11 at UNet batches 2 / 16 / 32 / 64."""
    from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine

    with plan_check.recording() as tr:
        for branch in range(12):
            tr.context = f"unet n_latents={n_latents} cfg={cfg} branch={branch}"
            tr.raws.clear()
            eng = UNetEngine(packed, n_latents=n_latents, cfg_dup=cfg, device="cpu", cache_branch=branch)
            assert set(eng.plans) == {"ctx", "full", "cached"}
        assert tr.n_ops > 12 * 380 and tr.arena_reads > 10000 and tr.arena_writes > 7000 and tr.extents_checked > 15000 and tr.overlaps_checked > 4000
        assert tr.problems == [], "\n".join(tr.problems[:10])


def test_vae_and_clip_plans_have_no_lifetime_hazards():
    from sonicdiffusionbayeslab_b200.clip_engine import ClipTextEngine, ClipVisionEngine
    from sonicdiffusionbayeslab_b200.metrics.metrics import make_clip_model
    from sonicdiffusionbayeslab_b200.text import make_text_encoder
    from sonicdiffusionbayeslab_b200.vae_engine import VaeEngine
    from sonicdiffusionbayeslab_b200.vae_spec import random_vae_state_dict

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clip, _ = make_clip_model(None)
        text = make_text_encoder(29, None)
    vae_sd = random_vae_state_dict(29)
    with plan_check.recording() as tr:
        for n_img, latent in ((1, 64), (2, 32), (16, 64)):
            tr.context = f"vae n_img={n_img} latent={latent}"
            tr.raws.clear()
            VaeEngine(dict(vae_sd), n_img=n_img, latent=latent, io_dtype=torch.bfloat16, device="cpu")
        n_vae = tr.n_ops
        c = text.config
        for n in (1, 32):
            for what, make in (("clip image tower", lambda: ClipVisionEngine(clip.state_dict(), n=n, device="cpu")),
                               ("clip text tower", lambda: ClipTextEngine(clip.state_dict(), n=n, device="cpu")),
                               ("prompt encoder", lambda: ClipTextEngine(
                                   text.state_dict(), n=n, seq=c.max_position_embeddings, width=c.hidden_size,
                                   heads=c.num_attention_heads, layers=c.num_hidden_layers, mlp=c.intermediate_size,
                                   device="cpu"))):
                tr.context = f"{what} n={n}"
                tr.raws.clear()                                       # one engine = one arena
                make()
        assert n_vae > 250 and tr.n_ops - n_vae > 300
        assert tr.problems == [], "\n".join(tr.problems[:10])


def test_recorded_plans_carry_the_algorithmic_flops_of_the_roofline(packed):
    """The roofline's numerator (bench.py ``FLOP_PER_SAMPLE_FWD``, SURVEY.md appendix B: 803.27 GFLOP per sample per UNet
    forward; 63.25 for the DeepCache branch-0 cached step) against the work the recorded plans really contain:
    2 M N K taps of every conv / linear launch (a folded LayerNorm's side chunk and the phase form of the upsample
    convolutions counted as the operator they implement, as csrc/gemm.cu does) + 4 B H Sq Sk d of every attention.
    The only excess is layout padding: conv_in reads 8 channels for 4, conv_out writes 16 for 4 (+0.05 %)."""
    import importlib.util

    from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine

    spec = importlib.util.spec_from_file_location("bench_for_constants", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    per_plan = {}
    with plan_check.recording() as tr:
        record = tr.on_op

        def on_op(name, args):
            h = args[0].value
            if name == "sonic_plan_add_conv_gemm":
                g = args[1]._obj
                k = g.c0 if g.row_scale else g.c0 + g.c1            # folded LayerNorm: a1 is the 64-column side tensor
                per_plan[h] = per_plan.get(h, 0.0) + 2.0 * g.n_img * g.H * g.W * (4 if g.upsample else 1) * g.N * k * g.taps
            elif name == "sonic_plan_add_attention":
                a = args[1]._obj
                per_plan[h] = per_plan.get(h, 0.0) + 4.0 * a.batch * a.heads * a.seq_q * a.seq_k * a.head_dim
            record(name, args)

        tr.on_op = on_op
        eng = UNetEngine(packed, n_latents=4, cfg_dup=True, device="cpu", cache_branch=0)
        flops = {name: per_plan.get(p.h.value, 0.0) for name, p in eng.plans.items()}
    samples = 8
    full = (flops["full"] + flops["ctx"]) / samples                  # the context K / V projections run once per call
    cached = flops["cached"] / samples
    assert abs(full / bench.FLOP_PER_SAMPLE_FWD - 1) < 1e-3, full
    assert 0 <= cached / 63.25e9 - 1 < 1e-2, cached


def test_cached_plans_read_exactly_the_resident_feature_of_the_full_plan(packed):
    """Data flow BETWEEN the three plans of an engine, for every DeepCache branch (the GPU parity
    tests run branches 0, 1, 2, 3 and 5): the full plan reads the 16 context K|V projections of the ctx plan; the cached
    plan reads from the full plan exactly TWO buffers -- the feature of the up block below the cut and the GroupNorm
    pre-reduction that travels with it -- and from the ctx plan the K|V of just the cross-attention layers it recomputes:
    with (block i, layer j) = divmod(branch, 3) those are the down layers above the cut (2 per attention block 0-2) and
    the up layers from the cut outwards (3 per attention block), i.e. 1, 3, 5 | 6, 8, 10 | 11, 13, 15 | 15, 15, 15."""
    from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine

    expected_ctx_reads = [1, 3, 5, 6, 8, 10, 11, 13, 15, 15, 15, 15]
    cached_ops = []
    with plan_check.recording() as tr:
        record = tr.on_op
        for branch in range(12):
            tr.raws.clear()
            writer, cross = {}, {}

            def base_of(p):
                return next((b for b, e in tr.raws.items() if b <= p < b + e[0]), None)

            def on_op(name, args):
                h = args[0].value
                ops = [(base_of(p), k) for p, k in tr._operands(name, args)]
                for b, k in ops:
                    if b is not None and k == "r" and writer.get(b, h) != h:
                        cross.setdefault((writer[b], h), set()).add(b)
                for b, k in ops:
                    if b is not None and k == "w":
                        writer[b] = h
                record(name, args)

            tr.on_op = on_op
            eng = UNetEngine(packed, n_latents=2, cfg_dup=True, device="cpu", cache_branch=branch)
            tr.on_op = record
            name_of = {p.h.value: n for n, p in eng.plans.items()}
            got = {(name_of[a], name_of[b]): len(v) for (a, b), v in cross.items()}
            assert got == {("ctx", "full"): 16, ("full", "cached"): 2, ("ctx", "cached"): expected_ctx_reads[branch]}, (
                branch, got)
            cached_ops.append(len(eng.plans["cached"].log))
        assert tr.problems == []
    assert cached_ops == sorted(cached_ops) and cached_ops[0] < 40 and cached_ops[-1] < 342   # deeper cut, more work


def test_plan_log_matches_the_operators_and_the_roofline_byte_counts_of_bench(packed):
    """bench.py pairs the native per-operator profile with ``plan.log`` line by line and parses the GroupNorm /
    ``ln_side`` lines for the HBM rooflines: one log line per recorded operator, in order, with the fields bench reads;
    at the benchmarked shape (UNet batch 32) the GroupNorm traffic is SURVEY 8(d)'s 180.2 MB per sample-forward."""
    from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine

    per_plan = {}
    with plan_check.recording() as tr:
        record = tr.on_op

        def on_op(name, args):
            per_plan.setdefault(args[0].value, []).append(name[len("sonic_plan_add_"):])
            record(name, args)

        tr.on_op = on_op
        eng = UNetEngine(packed, n_latents=16, cfg_dup=True, device="cpu")
        ops = {n: per_plan[p.h.value] for n, p in eng.plans.items()}
        logs = {n: list(p.log) for n, p in eng.plans.items()}
    gn_bytes = ln_sides = 0
    for name in ops:
        assert len(ops[name]) == len(logs[name]), name
        for op, text in zip(ops[name], logs[name]):
            head = text.split()[0]
            assert head.startswith(op.split("_fused")[0][:6]) or (op, head) in {("conv_gemm", "gemm"), ("conv_gemm", "conv3x3"),
                                                                               ("conv_gemm", "linear")}, (op, text)
            if name != "full":
                continue
            f = dict(p.split("=") for p in text.split()[1:] if "=" in p)        # bench.py's parse
            if op == "ln_side":
                ln_sides += 1
                assert int(f["rows"]) > 0 and int(f["parts"]) > 0
            elif op.startswith("groupnorm"):
                gn_bytes += 2 * int(f["rows"]) * int(f["C"]) * 2
    assert ln_sides == 48 and len(logs["full"]) == 342
    assert gn_bytes == 5767168000 and abs(gn_bytes / 32 / 180.2e6 - 1) < 1e-3
