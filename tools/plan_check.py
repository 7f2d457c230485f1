"""Buffer-lifetime check of the recorded launch plans, on the CPU.

The engines (unet_engine.py, vae_engine.py, clip_engine.py) record every kernel launch ONCE into a native plan and hand
activations out of an exact-size free-list arena; pointers are baked into the plan, so releasing a tensor one operator
too early is a silent corruption that only shows at the batch sizes where a later allocation happens to have the same
rounded size.  Recording touches no device memory, so the whole recording can be replayed here with the library
replaced by a recorder (``sonic_gemm_block_n`` -- a host function -- stays the real one) and checked operator by
operator, in execution order, for every plan variant (batch size, guidance on / off, every DeepCache branch):

  1. every pointer operand that lies in an arena buffer is LIVE (allocated, not released) when its operator is recorded;
  2. every arena buffer an operator READS (``const`` in include/sonic.h) was WRITTEN (non-``const``) by an earlier
     operator since that buffer was last handed out -- i.e. nobody reads a recycled buffer through a stale pointer,
     and the cached (DeepCache) plan reads only features the full plan left resident;
  3. the arena never hands out a live buffer and nothing is released twice;
  4. the bytes an operator touches -- from its documented semantics: rows x pitch of every GEMM / attention / norm
     operand, the GroupNorm / LayerNorm partial-statistics tables -- stay inside the arena buffer they start in;
  5. a GEMM / attention output never overlaps one of its own inputs (the element-wise residual may);
  6. every TMA operand starts on a 16-byte boundary of its buffer and has a pitch that is a multiple of 16 bytes.

    python tools/plan_check.py            # UNet batches 2 ... 64, guidance on / off, DeepCache branches 0-11, VAE, CLIP
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _header_constness():
    """{function: [None | 'r' | 'w' per parameter]} and {struct: {field: 'r' | 'w'}} from include/sonic.h: a pointer to
    const is read, any other pointer is written (or updated in place)."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "sonic.h")).read(), flags=re.S)

    def kind(decl):
        decl = " ".join(decl.split())
        if "*" not in decl:
            return None
        return "r" if decl.startswith("const") else "w"

    funcs = {}
    for m in re.finditer(r"\bint\s+(sonic_plan_add_\w+)\s*\(([^;{]*?)\)\s*;", hdr, re.S):
        funcs[m.group(1)] = [kind(p) for p in m.group(2).split(",")]
    structs = {}
    for name in ("sonic_gemm_args", "sonic_attention_args"):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        fields = {}
        for decl in body.split(";"):
            decl = decl.strip()
            if "*" in decl:
                first = decl.split(",")[0]
                fields[first.split()[-1].lstrip("*")] = kind(first)
        structs[name] = fields
    return funcs, structs


class Tracker:
    def __init__(self):
        self.funcs, self.structs = _header_constness()
        self.raws = {}                   # base pointer -> [nbytes, live, generation, generation of the last write]
        self.n_ops = 0
        self.arena_reads = self.arena_writes = self.extents_checked = self.overlaps_checked = 0   # operands the checks really looked at
        self.problems = []
        self.context = ""

    # ---- arena events
    def on_alloc(self, raw):
        e = self.raws.setdefault(raw.data_ptr(), [raw.numel(), False, 0, -1])
        if e[1]:
            self.problems.append(f"{self.context}: the arena handed out a live buffer")
        e[1] = True
        e[2] += 1

    def on_release(self, raw):
        e = self.raws.get(raw.data_ptr())
        if e is None or not e[1]:
            self.problems.append(f"{self.context}: buffer released twice (or never allocated)")
            return
        e[1] = False

    # ---- operator events
    def _find(self, p):
        for base, e in self.raws.items():
            if base <= p < base + e[0]:
                return e
        return None

    def _operands(self, name, args):
        """[(pointer value, 'r' | 'w')] of one recorded call (args[0] is the plan handle)."""
        out = []
        kinds = self.funcs.get(name)
        if kinds is None:
            raise KeyError(f"{name} is not declared in include/sonic.h")
        if len(kinds) != len(args):
            raise TypeError(f"{name}: {len(args)} arguments, the header declares {len(kinds)}")
        for a, k in zip(args[1:], kinds[1:]):
            if hasattr(a, "_obj"):                                    # byref(struct)
                st = a._obj
                fields = self.structs["sonic_gemm_args" if type(st).__name__ == "GemmArgs" else "sonic_attention_args"]
                for f, tp in st._fields_:
                    if tp is C.c_void_p and getattr(st, f):
                        out.append((getattr(st, f), fields[f]))
            elif isinstance(a, C.Array):                              # host table of device pointers (gemv)
                if k is not None and a._type_ in (C.c_void_p,):
                    out += [(v, k) for v in a if v]
            elif isinstance(a, C.c_void_p):
                if a.value and k is not None:
                    out.append((a.value, k))
        return out

    def on_op(self, name, args):
        self.n_ops += 1
        where = f"{self.context} op {self.n_ops} {name}"
        ops = self._operands(name, args)
        for p, k in ops:                                              # reads first: in-place operators read, then write
            e = self._find(p)
            if e is None:
                continue                                              # weights, inputs, persistent buffers
            self.arena_reads += k == "r"
            self.arena_writes += k == "w"
            if not e[1]:
                self.problems.append(f"{where}: operand {p:#x} ({k}) lies in a RELEASED arena buffer")
            elif k == "r" and e[3] != e[2]:
                self.problems.append(f"{where}: reads arena buffer {p:#x} that nobody has written since it was handed out")
        for p, k in ops:
            e = self._find(p)
            if e is not None and k == "w":
                e[3] = e[2]
        for what, p, nbytes in self._extents(name, args):             # 4. operands stay inside their arena buffer
            if not p:
                continue
            for base, e in self.raws.items():
                if base <= p < base + e[0]:
                    self.extents_checked += 1
                    if p + nbytes > base + e[0]:
                        self.problems.append(f"{where}: {what} needs {nbytes} bytes at {p:#x}, its arena buffer ends "
                                             f"{base + e[0] - p} bytes after that")
                    break

        self._aliasing_and_alignment(name, args, where)

    def _aliasing_and_alignment(self, name, args, where):
        """5. a GEMM / attention output never overlaps one of its own inputs (other CTAs still read them; the residual,
        added element-wise by the CTA that writes the same element, is the one in-place operand); 6. every TMA operand
        starts on a 16-byte boundary of its buffer with a pitch that is a multiple of 16 bytes."""
        ext = {w: (p, n) for w, p, n in self._extents(name, args) if p}
        if name == "sonic_plan_add_conv_gemm":
            outs, ins = ["out"], ["a0", "a1"]
            g = args[1]._obj
            pitches = {"a0": g.ld0, "a1": g.ld1, "out": g.ld_out, "residual": g.ld_res}
        elif name == "sonic_plan_add_attention":
            outs, ins = ["o"], ["q", "k", "v"]
            a = args[1]._obj
            pitches = {"q": a.ld_q, "k": a.ld_k, "v": a.ld_v, "o": a.ld_o}
        else:
            return
        for o in outs:
            for i in ins:
                if o in ext and i in ext:
                    (po, no), (pi, ni) = ext[o], ext[i]
                    self.overlaps_checked += 1
                    if po < pi + ni and pi < po + no:
                        self.problems.append(f"{where}: output `{o}` [{po:#x}, +{no}) overlaps input `{i}` [{pi:#x}, +{ni})")
        for w, ld in pitches.items():
            if w not in ext:
                continue
            p = ext[w][0]
            base = next((b for b, e in self.raws.items() if b <= p < b + e[0]), None)
            off = p - base if base is not None else p          # non-arena tensors: torch allocations are 64-byte aligned
            if off % 16 or (ld * 2) % 16:
                self.problems.append(f"{where}: `{w}` is not TMA-addressable (offset {off} bytes, pitch {ld} elements)")

    @staticmethod
    def _extents(name, args):
        """[(operand, pointer, bytes the kernel touches)] from the operator semantics documented in include/sonic.h."""
        def v(a):
            return a.value if isinstance(a, C.c_void_p) else (a.value if hasattr(a, "value") else a)

        out = []
        if name == "sonic_plan_add_conv_gemm":
            g = args[1]._obj
            m_src = g.n_img * g.H * g.W                                # rows of the launch's pixel grid
            rows_in = m_src * (4 if g.stride == 2 else 1)              # stride 2: H, W are the OUTPUT extents
            rows_out = m_src * (4 if g.upsample else 1)                # upsample: H, W are the SOURCE extents
            width = g.N // 2 if g.epilogue == 1 else g.N               # GEGLU halves the columns
            out.append(("a0", g.a0, ((rows_in - 1) * g.ld0 + g.c0) * 2))
            if g.a1:
                out.append(("a1", g.a1, ((rows_in - 1) * g.ld1 + g.c1) * 2))
            out.append(("out", g.out, ((rows_out - 1) * g.ld_out + width) * 2))
            if g.residual:
                out.append(("residual", g.residual, ((rows_out - 1) * g.ld_res + width) * 2))
            if g.row_scale:
                out.append(("row_scale", g.row_scale, rows_out * 4))
            if g.gn_partial and not g.upsample:
                out.append(("gn_partial", g.gn_partial, -(-rows_out // 32) * width * 2 * 4))
            if g.ln_stats_out and g.block_n:
                out.append(("ln_stats_out", g.ln_stats_out, rows_out * 2 * -(-g.N // g.block_n) * 2 * 4))
        elif name == "sonic_plan_add_attention":
            a = args[1]._obj
            hd = a.heads * a.head_dim
            out += [("q", a.q, ((a.batch * a.seq_q - 1) * a.ld_q + hd) * 2), ("k", a.k, ((a.batch * a.seq_k - 1) * a.ld_k + hd) * 2),
                    ("v", a.v, ((a.batch * a.seq_k - 1) * a.ld_v + hd) * 2), ("o", a.o, ((a.batch * a.seq_q - 1) * a.ld_o + hd) * 2)]
        elif name == "sonic_plan_add_layernorm":
            _, x, y, rows, c = args[:5]
            out += [("x", v(x), rows * c * 2), ("y", v(y), rows * c * 2)]
        elif name in ("sonic_plan_add_groupnorm", "sonic_plan_add_groupnorm_fused"):
            if name.endswith("fused"):
                _, x0, c0, p0, x1, c1, p1, n_img, hw = args[:9]
                y = args[-1]
                blocks = -(-(n_img * hw) // 32)
                out += [("part0", v(p0), blocks * c0 * 2 * 4), ("part1", v(p1), blocks * c1 * 2 * 4)]
            else:
                _, x0, c0, x1, c1, n_img, hw = args[:7]
                y = args[-1]
            out += [("x0", v(x0), n_img * hw * c0 * 2), ("x1", v(x1), n_img * hw * c1 * 2),
                    ("y", v(y), n_img * hw * (c0 + c1) * 2)]
        elif name == "sonic_plan_add_ln_side":
            _, partials, parts, m, _k, _eps, side, rstd = args
            out += [("partials", v(partials), m * parts * 2 * 4), ("side", v(side), m * 64 * 2), ("rstd", v(rstd), m * 4)]
        elif name == "sonic_plan_add_softmax_rows":
            _, x, rows, cols, ld, _scale = args
            out.append(("x", v(x), ((rows - 1) * v(ld) + cols) * 2))
        return out


@contextlib.contextmanager
def recording():
    """Replace the library and the arena hooks inside the engine modules; yields the Tracker."""
    import torch

    from sonicdiffusionbayeslab_b200 import _lib, clip_engine, kernels, unet_engine, vae_engine

    real = _lib.lib()
    tracker = Tracker()
    handles = [0]

    class Recorder:
        def __getattr__(self, name):
            if name in ("sonic_gemm_block_n", "sonic_version", "sonic_last_error"):
                return getattr(real, name)

            def call(*args):
                if name == "sonic_plan_create":
                    handles[0] += 1
                    args[0]._obj.value = handles[0]
                elif name.startswith("sonic_plan_add_"):
                    tracker.on_op(name, args)
                elif name not in ("sonic_plan_destroy",):
                    raise RuntimeError(f"{name} needs a GPU: the check only records plans")
                return 0

            return call

    fake = Recorder()
    mods = [m for m in (_lib, kernels, unet_engine, vae_engine, clip_engine) if hasattr(m, "lib")]
    saved = [(m, m.lib) for m in mods]
    arena = unet_engine.Arena
    saved_arena = (arena.alloc, arena.release)
    saved_cuda = torch.Tensor.is_cuda

    def alloc(self, shape, dtype=torch.bfloat16):
        t = saved_arena[0](self, shape, dtype)
        tracker.on_alloc(t._arena_raw)
        return t

    def release(self, t):
        for attr in ("_gn_part", "_ln_part"):                         # as Arena.release: partial buffers travel along
            part = getattr(t, attr, None)
            if part is not None:
                setattr(t, attr, None)
                release(self, part)
        tracker.on_release(t._arena_raw)
        self.free.setdefault(t._arena_raw.numel(), []).append(t._arena_raw)

    import weakref

    plan_cls = unet_engine._Plan
    saved_init, plans = plan_cls.__init__, []

    def plan_init(self, *a, **k):
        saved_init(self, *a, **k)
        plans.append(weakref.ref(self))

    try:
        for m in mods:
            m.lib = lambda: fake
        arena.alloc, arena.release = alloc, release
        plan_cls.__init__ = plan_init
        torch.Tensor.is_cuda = property(lambda self: True)             # host tensors stand in for device tensors
        yield tracker
    finally:
        for ref in plans:                                             # recorded plans hold recorder handles: a plan that
            plan = ref()                                              # outlives this block must not reach the real
            if plan is not None:                                      # sonic_plan_destroy from its __del__
                plan.h = C.c_void_p()
        for m, f in saved:
            m.lib = f
        arena.alloc, arena.release = saved_arena
        plan_cls.__init__ = saved_init
        torch.Tensor.is_cuda = saved_cuda


def main():
    import torch

    from sonicdiffusionbayeslab_b200.unet_engine import PackedWeights, UNetEngine
    from sonicdiffusionbayeslab_b200.unet_spec import random_unet_state_dict

    total = 0
    with recording() as tr:
        w = PackedWeights(random_unet_state_dict(29), "cpu")
        for n_lat in (1, 2, 8, 16, 32):
            for cfg in (True, False):
                for branch in range(12):
                    tr.context = f"unet n_latents={n_lat} cfg={cfg} branch={branch}"
                    tr.raws.clear()
                    before = tr.n_ops
                    UNetEngine(w, n_latents=n_lat, cfg_dup=cfg, device="cpu", cache_branch=branch)
                    total += tr.n_ops - before
        print(f"UNet: {total} operators recorded over 120 plan sets ({tr.arena_reads} arena reads, {tr.arena_writes} "
              f"arena writes, {tr.extents_checked} operand extents checked), {len(tr.problems)} problems")
        from sonicdiffusionbayeslab_b200.clip_engine import ClipTextEngine, ClipVisionEngine
        from sonicdiffusionbayeslab_b200.metrics.metrics import make_clip_model
        from sonicdiffusionbayeslab_b200.text import make_text_encoder
        from sonicdiffusionbayeslab_b200.vae_engine import VaeEngine
        from sonicdiffusionbayeslab_b200.vae_spec import random_vae_state_dict

        before = tr.n_ops
        vae_sd = random_vae_state_dict(29)
        for n_img, latent in ((1, 64), (2, 32), (4, 64), (16, 64)):
            tr.context = f"vae n_img={n_img} latent={latent}"
            tr.raws.clear()
            VaeEngine(dict(vae_sd), n_img=n_img, latent=latent, io_dtype=torch.bfloat16, device="cpu")
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            clip, _ = make_clip_model(None)
            text = make_text_encoder(29, None)
        sd = clip.state_dict()
        for n in (1, 4, 32):
            # one engine = one arena: the table of arena buffers is per engine (a dead engine's addresses are recycled
            # by the host allocator for anything, weights included)
            tr.context = f"clip image tower n={n}"
            tr.raws.clear()
            ClipVisionEngine(sd, n=n, device="cpu")
            tr.context = f"clip text tower n={n}"
            tr.raws.clear()
            ClipTextEngine(sd, n=n, device="cpu")
            tr.context = f"prompt encoder n={n}"
            tr.raws.clear()
            c = text.config
            ClipTextEngine(text.state_dict(), n=n, seq=c.max_position_embeddings, width=c.hidden_size,
                           heads=c.num_attention_heads, layers=c.num_hidden_layers, mlp=c.intermediate_size, device="cpu")
        print(f"VAE decoder / CLIP towers / prompt encoder: {tr.n_ops - before} operators, {len(tr.problems)} problems "
              "in total")
        for p in tr.problems[:20]:
            print("  ", p)
    return 1 if tr.problems else 0


if __name__ == "__main__":
    sys.exit(main())
